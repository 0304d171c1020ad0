"""Test helper: a recording stand-in for matplotlib (absent from this image) so that the plot methods of the drop-in
MonteCarloAnalyzer (reference monte_carlo.py:562-707) can be exercised: every Axes / pyplot call is logged with its
arguments, nothing is drawn.  `install()` puts it into sys.modules; `remove()` takes it out again."""
import sys
import types

CALLS = []


class _Recorder:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, attr):
        def call(*args, **kwargs):
            CALLS.append((self._name, attr, args, kwargs))
            if attr == "add_subplot":
                return _Recorder("axes3d" if kwargs.get("projection") == "3d" else "axes")
            return None
        return call


class _Grid:
    def __init__(self, rows, cols):
        self._a = [[_Recorder(f"axes[{r},{c}]") for c in range(cols)] for r in range(rows)]
        self._rows, self._cols = rows, cols

    def __getitem__(self, key):
        r, c = key
        return self._a[r][c]

    def __iter__(self):
        return iter(self._a[0] if self._rows == 1 else self._a)


def _subplots(rows=1, cols=1, **kwargs):
    CALLS.append(("pyplot", "subplots", (rows, cols), kwargs))
    if rows == 1 and cols == 1:
        return _Recorder("figure"), _Recorder("axes")
    g = _Grid(rows, cols)
    return _Recorder("figure"), (tuple(g._a[0]) if rows == 1 else g)


def _savefig(path, **kwargs):
    CALLS.append(("pyplot", "savefig", (path,), kwargs))
    with open(path, "wb") as fh:
        fh.write(b"stub")


def install():
    del CALLS[:]
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    plt.subplots = _subplots
    plt.figure = lambda **k: _Recorder("figure")
    plt.tight_layout = lambda *a, **k: None
    plt.savefig = _savefig
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    return CALLS


def remove():
    sys.modules.pop("matplotlib", None)
    sys.modules.pop("matplotlib.pyplot", None)
