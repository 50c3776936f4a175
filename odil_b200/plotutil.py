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

_extlist = None
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


def set_extlist(extlist=None):
    """File extensions `savefig` writes; None re-reads ODIL_EXTLIST (comma-separated, default png)."""
    global _extlist
    _extlist = os.environ.get("ODIL_EXTLIST", "png").split(",") if extlist is None else extlist


set_extlist()


def apply_clip_box(ax, artists, lower=(0, 0), upper=(1, 1.02)):
    """Clips `artists` to the box [lower, upper] given in axes coordinates of `ax`."""
    tr = _matplotlib().transforms
    box = tr.TransformedBbox(tr.Bbox([lower, upper]), ax.transAxes)
    for a in artists:
        a.set_clip_box(box)


# time stamps are blanked so that re-running a script reproduces the files bit for bit
_NO_DATES = {"svg": {"Date": None}, "pdf": {"DateModified": None, "CreationDate": None}}


def savefig(fig, path_without_ext, extlist=None, skip_existing=False, printf=None, **kwargs):
    """Saves `fig` once per extension (`extlist`, default from `set_extlist`); `printf` receives each path."""
    say = printf if printf is not None else (lambda _: None)
    for ext in (_extlist if extlist is None else extlist):
        path = path_without_ext + "." + ext
        if skip_existing and os.path.isfile(path):
            say("skip existing '{}'".format(path))
            continue
        say(path)
        fig.savefig(path, metadata=_NO_DATES.get(ext, {}), **kwargs)


def savelegend(fig, ax, path, **kwargs):
    """Saves the legend of `ax` alone, cropped to its extent."""
    _matplotlib()
    import matplotlib.pyplot as plt

    figleg, axleg = plt.subplots()
    handles, labels = ax.get_legend_handles_labels()
    legend = axleg.legend(handles, labels, loc="center", frameon=False)
    axleg.set_axis_off()
    figleg.canvas.draw()
    bbox = legend.get_window_extent().transformed(fig.dpi_scale_trans.inverted())
    savefig(figleg, path, bbox_inches=bbox, **kwargs)


def set_log_ticks(axis):
    """Unlabelled minor ticks at 2..9 x 10^k on a logarithmic axis."""
    import numpy as np

    ticker = _matplotlib().ticker
    axis.set_minor_locator(ticker.LogLocator(base=10.0, subs=np.arange(0.1, 0.99, 0.1), numticks=12))
    axis.set_minor_formatter(ticker.NullFormatter())
