class Obstacles:
    """Ground circle record, as robot_models/obstacles.py:6-10 in the reference."""

    def __init__(self, x, y, radius):
        self.x, self.y, self.radius = x, y, radius
