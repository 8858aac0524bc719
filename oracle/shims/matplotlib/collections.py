"""LineCollection stand-in (see matplotlib/__init__.py): keeps what rasterize_forest passes (tree2img.py:103)."""


class LineCollection:
    def __init__(self, segments, linewidths=None, colors=None, antialiaseds=None, capstyle=None, **kw):
        if not (capstyle == "round" and antialiaseds is True and isinstance(colors, str) and colors == "w"):
            raise NotImplementedError("the matplotlib stand-in only renders white, anti-aliased, round-capped collections")
        self.segments = segments
        self.linewidths = linewidths
