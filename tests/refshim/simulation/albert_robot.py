"""physical_sim=True is out of scope (BASELINE configs: "no pybullet").  TEST INFRASTRUCTURE."""


def setup_environment(*a, **k):
    raise RuntimeError("pybullet simulation is not available: run the Interface with physical_sim=False")


def run_step(*a, **k):
    raise RuntimeError("pybullet simulation is not available: run the Interface with physical_sim=False")
