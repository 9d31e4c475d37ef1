"""`from CNs_backup.backups import CN_forecast as CN` -> the B200 drop-in module."""
import ComplexNetworks as CN_forecast  # noqa: F401  (resolved through compat/ComplexNetworks.py)
