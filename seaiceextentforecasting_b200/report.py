"""Output layout of the retrospective scripts: the two CSV tables
`<Init>1st_detrended_forecasts_<fmin>-<fmax>.csv` and `<Init>1st_forecasts_with_trend_<fmin>-<fmax>.csv`
(north/retrospective_forecasts/June1st_retro.py:344-364; every sibling script writes the same layout with its own
init name and, in the south, region names).  Host-only bookkeeping over the arrays `SweepPlan.assemble()` returns."""
from __future__ import annotations

import numpy as np

FIRST_YEAR = 1979
INIT_NAME = {"north_june": "June1st", "north_july": "July1st", "north_august": "August1st",
             "north_september": "September1st", "south_february": "February1st", "south_january": "January1st",
             "south_december": "December1st"}


def _prep(data, skill=None):
    data = list(np.asarray(data).tolist()) if not isinstance(data, list) else list(data)
    data.append(skill if skill is not None else '')
    return data


def retro_tables(plan, gpr, cfg_name):
    """-> (df_dt, df_rt): pandas DataFrames with the index / columns / trailing `Skill` row of the reference
    (June1st_retro.py:344-361).  `gpr` = SweepPlan.assemble(...)."""
    import pandas as pd
    cfg = next(c for c in plan.cfgs if c.name == cfg_name)
    g = gpr[cfg_name]
    skill_rt, skill_dt = plan.skill(gpr)[cfg_name]
    fmin, fmax = plan.fmin, plan.fmax
    years = np.arange(fmin, fmax + 1).tolist()
    years.append('Skill')
    regs = cfg.regions
    columns1, columns2, cols_dt, cols_rt = [], [], [], []
    for k, reg in enumerate(regs):
        columns1 += [reg + '$_o$', reg + '$_f$', reg + '$_f$ unc']
        columns2 += [reg + '$_o$', reg + '$_f$']
        dt_obs = [plan.sie_dt[reg][t - (fmin - 1), t - FIRST_YEAR] for t in range(fmin, fmax + 1)]
        cols_dt += [_prep(dt_obs), _prep(g[reg + '_fmean'], skill_dt[k]), _prep(np.sqrt(g[reg + '_fvar']).round(3))]
        cols_rt += [_prep(plan.sie[reg][fmin - FIRST_YEAR:fmax - FIRST_YEAR + 1]), _prep(g[reg + '_fmean_rt'], skill_rt[k])]
    df_dt = pd.DataFrame(list(zip(*cols_dt)), index=years, columns=columns1)
    df_rt = pd.DataFrame(list(zip(*cols_rt)), index=years, columns=columns2)
    return df_dt, df_rt


def write_csv(plan, gpr, cfg_name, directory="."):
    """Writes the two files with the reference's names (June1st_retro.py:363-364); returns their paths."""
    import os
    df_dt, df_rt = retro_tables(plan, gpr, cfg_name)
    init = INIT_NAME[cfg_name]
    span = f"{plan.fmin}-{plan.fmax}"
    p1 = os.path.join(directory, f"{init}_detrended_forecasts_{span}.csv")
    p2 = os.path.join(directory, f"{init}_forecasts_with_trend_{span}.csv")
    df_dt.to_csv(p1)
    df_rt.to_csv(p2)
    return p1, p2
