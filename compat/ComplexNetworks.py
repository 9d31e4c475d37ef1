"""Import shim: put this directory on sys.path and the reference scripts' own import lines resolve to the B200
implementation -- `import ComplexNetworks as CN` (north/September1st.py:160, all south scripts) and
`from ComplexNetworks import CN` (north/June1st.py:197, July1st.py:159, August1st.py:159)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from seaiceextentforecasting_b200.ComplexNetworks import Network  # noqa: E402,F401

CN = sys.modules[__name__]
