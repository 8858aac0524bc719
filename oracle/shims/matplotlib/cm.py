"""stub (see matplotlib/__init__.py): colour maps are only used by `colorize`, which the stand-in does not render"""


def plasma(values):
    raise NotImplementedError("colorize needs real matplotlib")
