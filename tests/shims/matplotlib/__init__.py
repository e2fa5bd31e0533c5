"""Test-only stand-in for matplotlib (absent, no network): os2d/utils/visualization.py imports matplotlib.pyplot at module level;
no plotting function is reached with the visualisation switches of the config off."""


def use(*args, **kwargs):
    pass
