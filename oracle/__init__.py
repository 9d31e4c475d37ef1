"""ORACLE -- test infrastructure, NOT product code.

CPU restatement of the reference's hot path (ComplexNetworks.Network + the script-level detrend/forecast/MLII),
delegating arithmetic to the same numpy/scipy routines the reference calls.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product
package `seaiceextentforecasting_b200` never does (tests/test_boundary.py greps for that).
Parity pinning: tests/golden/*.npz are outputs of the *unmodified reference* (imported from /root/reference
in the authoring container by tests/golden/make_golden.py); tests/test_oracle_golden.py checks this package
against them."""
