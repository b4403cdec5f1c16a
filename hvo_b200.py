"""Import alias: the package directory carries the repository's long hyphenated name, which the `import`
statement cannot spell.  `import hvo_b200` gives the same module object."""
import importlib
import sys

_pkg = importlib.import_module('a-low-texture-robust-hybrid-feature-based-visual-odometry_b200')
sys.modules[__name__] = _pkg
