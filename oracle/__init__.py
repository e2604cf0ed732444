"""CPU oracle for the B200 text->mel->waveform path.  TEST INFRASTRUCTURE ONLY.

This package restates, as flat functions over a state_dict, the arithmetic of the reference
modules on the hot path (each function cites the reference file:line it follows). It is pinned
against the real reference code run in-process (`oracle/make_golden.py` -> `tests/golden/*.npz`);
the reference itself ships no tests or golden vectors, so those fixtures are the
pin. `make_ref.py` copies the unmodified reference modules of the path to `oracle/_ref/` (git-ignored;
`ref_runner.py` runs them): that is what `bench.py` times as the reference arm and as the CPU / eager-GPU
baselines. Only `tests/`, `__graft_entry__.smoke()` / `build()` and `bench.py`'s baseline legs may import
anything under `oracle/`; the product package `tts_arabic_pytorch_b200` never does.
"""
