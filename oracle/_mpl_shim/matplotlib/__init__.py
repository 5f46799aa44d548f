"""Stand-in for matplotlib (absent from this image) used ONLY while importing the
reference from /root/reference to generate golden vectors.  Test infrastructure."""
import sys, types
from unittest.mock import MagicMock


class _Shim(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


def use(*_a, **_k):
    return None


pyplot = _Shim("matplotlib.pyplot")
pyplot.subplots = lambda *a, **k: (MagicMock(), MagicMock())
animation = _Shim("matplotlib.animation")
cm = _Shim("matplotlib.cm")
colors = _Shim("matplotlib.colors")
for _m in (pyplot, animation, cm, colors):
    sys.modules[_m.__name__] = _m
