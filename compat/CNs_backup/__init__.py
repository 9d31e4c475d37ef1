"""Alias package for `from CNs_backup.backups import CN_forecast as CN`
(north/retrospective_forecasts/June1st_retro.py:198 and siblings import a module the reference repo does not ship)."""
