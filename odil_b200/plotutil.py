"""
Figure helpers the reference's example scripts call (`from odil import plotutil`; reference src/odil/plotutil.py:
`set_extlist` :21-29, `apply_clip_box` :32-35, `savefig` :38-66, `savelegend` :69-76, `set_log_ticks` :79-82).
Plotting is not on the B200 hot path; this module exists so that problem scripts written for the reference import
and run unchanged.  matplotlib is imported on first use, so importing the module costs nothing and works on hosts
without matplotlib (the GPU boxes); a figure call without it raises an ImportError that says so.

Environment, as upstream: ODIL_AGG (default 1: non-interactive Agg canvas), ODIL_EXTLIST (default "png").
The reference's house style sheet (`odil.mplstyle`) is a matter of looks and is not shipped.
"""
import os

_mpl = None


def _matplotlib():
    """matplotlib, configured once (Agg canvas unless ODIL_AGG=0)."""
    global _mpl
    if _mpl is None:
        try:
            import matplotlib
        except ImportError as e:
            raise ImportError("odil.plotutil needs matplotlib for figure output; it is not installed here "
                              "(run with plotting disabled, e.g. --plot 0 --plot_every 0)") from e
        if int(os.environ.get("ODIL_AGG", 1)):
            matplotlib.use("Agg")
        import logging

        logging.getLogger("matplotlib.font_manager").setLevel(logging.ERROR)
        _mpl = matplotlib
    return _mpl


class _Settings:
    extensions = ["png"]


def set_extlist(extlist=None):
    """Chooses the file types `savefig` writes: a list of extensions, or None to take them from the environment
    variable ODIL_EXTLIST (comma-separated, "png" when unset)."""
    _Settings.extensions = list(extlist) if extlist is not None else os.environ.get("ODIL_EXTLIST", "png").split(",")


set_extlist()


def apply_clip_box(ax, artists, lower=(0, 0), upper=(1, 1.02)):
    """Restricts drawing of `artists` to the rectangle lower..upper, given as fractions of the axes `ax`."""
    transforms = _matplotlib().transforms
    region = transforms.TransformedBbox(transforms.Bbox.from_extents(*lower, *upper), ax.transAxes)
    for artist in artists:
        artist.set_clip_box(region)


def _without_timestamps(ext):
    """Metadata that blanks the creation dates vector formats embed, so that reruns give identical files."""
    return {"svg": {"Date": None}, "pdf": {"CreationDate": None, "DateModified": None}}.get(ext, {})


def savefig(fig, path_without_ext, extlist=None, skip_existing=False, printf=None, **kwargs):
    """Writes `fig` to `path_without_ext.<ext>` for every extension (default: `set_extlist`).  Existing files are
    kept when `skip_existing`; `printf`, if given, is told each path (or that it was skipped)."""
    for ext in (_Settings.extensions if extlist is None else extlist):
        target = "{}.{}".format(path_without_ext, ext)
        exists = skip_existing and os.path.isfile(target)
        if printf is not None:
            printf("skip existing '{}'".format(target) if exists else target)
        if not exists:
            fig.savefig(target, metadata=_without_timestamps(ext), **kwargs)


def savelegend(fig, ax, path, **kwargs):
    """Writes the legend of `ax` as a figure of its own, cropped to the legend."""
    _matplotlib()
    from matplotlib import pyplot

    sheet, blank = pyplot.subplots()
    blank.set_axis_off()
    entries = blank.legend(*ax.get_legend_handles_labels(), loc="center", frameon=False)
    sheet.canvas.draw()
    extent = entries.get_window_extent().transformed(fig.dpi_scale_trans.inverted())
    savefig(sheet, path, bbox_inches=extent, **kwargs)


def set_log_ticks(axis):
    """Minor ticks without labels at 2, 3, ..., 9 times every power of ten of a logarithmic axis."""
    ticker = _matplotlib().ticker
    axis.set_minor_locator(ticker.LogLocator(base=10.0, subs=[k / 10 for k in range(1, 10)], numticks=12))
    axis.set_minor_formatter(ticker.NullFormatter())
