"""Per-script GP configuration of the reference (it has no config system: every value below is hard-coded
at the top of each script's `forecast()`).

    ls = np.logspace(-7, 2, 20),  ss = np.logspace(-3, 9, 20)

| init            | cite                                   |
| N June   (May)  | north/June1st.py:210-227               | SIC r>0, SST r<0 negated, z-score
| N July   (June) | north/July1st.py:169-183               | r>0
| N August (July) | north/August1st.py:169-186             | region 0 all; others r>0 & p/2<0.08
| N Sept.  (Aug)  | north/September1st.py:170-187          | region 0 all; others r>0 & p/2<0.05
| S Feb.   (Jan)  | south/February1st.py:162-178           | region 0 all; others r>0 & p/2<0.05
| S Jan.   (Dec)  | south/January1st.py:163-179            | region 0 all; others r>0 & p/2<0.08; previous-year network
| S Dec.   (Nov)  | south/December1st.py:162-176           | r>0; previous-year network
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

LS = np.logspace(-7, 2, 20)
SS = np.logspace(-3, 9, 20)

RULE_POS, RULE_ALL, RULE_POS_SIG = 0, 1, 2


@dataclass(frozen=True)
class ForecastConfig:
    name: str
    hemisphere: str                     # 'north' | 'south'
    regions: tuple
    ell: tuple                          # l per region
    sig: tuple                          # sigma_n tilde per region
    rule: tuple                         # per region: RULE_*
    alpha: float = 0.05                 # for RULE_POS_SIG
    zscore: bool = False
    use_sst: bool = False
    prev_year_network: bool = False     # south January/December: network of year-1, y drops 1979


_N = ("Pan-Arctic", "Beaufort", "Chukchi")
_S = ("Pan-Antarctic", "Ross", "Weddell")

CONFIGS = {
    "north_june": ForecastConfig("north_june", "north", _N, (LS[16], LS[14], LS[12]), (SS[1], SS[4], SS[6]),
                                 (RULE_POS,) * 3, zscore=True, use_sst=True),
    "north_july": ForecastConfig("north_july", "north", _N, (LS[11], LS[0], 3.125433e+10),
                                 (SS[4], SS[15], 40221.26298973), (RULE_POS,) * 3),
    "north_august": ForecastConfig("north_august", "north", _N, (LS[9], LS[7], LS[3]), (SS[4], SS[13], SS[13]),
                                   (RULE_ALL, RULE_POS_SIG, RULE_POS_SIG), alpha=0.08),
    "north_september": ForecastConfig("north_september", "north", _N, (LS[8], LS[9], LS[3]),
                                      (SS[6], SS[3], SS[13]), (RULE_ALL, RULE_POS_SIG, RULE_POS_SIG), alpha=0.05),
    "south_february": ForecastConfig("south_february", "south", _S, (LS[16], LS[5], LS[3]), (SS[0], SS[11], SS[13]),
                                     (RULE_ALL, RULE_POS_SIG, RULE_POS_SIG), alpha=0.05),
    "south_january": ForecastConfig("south_january", "south", _S, (LS[2], LS[1], LS[3]), (SS[14],) * 3,
                                    (RULE_ALL, RULE_POS_SIG, RULE_POS_SIG), alpha=0.08, prev_year_network=True),
    "south_december": ForecastConfig("south_december", "south", _S, (LS[4], LS[9], LS[2]), (SS[13], SS[4], SS[13]),
                                     (RULE_POS,) * 3, prev_year_network=True),
}

NORTH_INITS = ("north_june", "north_july", "north_august", "north_september")
