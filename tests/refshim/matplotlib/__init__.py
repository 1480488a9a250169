"""Empty stand-in: the reference's hot-path modules import matplotlib at module top (plots only).  TEST INFRASTRUCTURE."""
