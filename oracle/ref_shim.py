"""Run the UNMODIFIED reference (`/root/reference`) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing on the product path imports this file; it is
used (a) by `oracle/make_golden.py` to freeze golden fixtures under
`tests/golden/`, and (b) by the `-m "not gpu"` tests that pin the C
restatement (`oracle/oracle_seir.c`) against the reference when
`/root/reference` is present (it does not exist on the GPU box).

Recipe follows SURVEY.md section 8c:
  * numpy >= 1.24 removed `np.int` / `np.bool8` which the reference uses
    (`sample_households.py:40`, `seir_individual.py:165`): inject the aliases
    before importing the reference modules (no reference source is edited);
  * the reference opens data files relative to the CWD and writes its population
    pickle there, and `/root/reference` is read-only: run from a scratch
    directory holding symlinks to the data files.
"""
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("SEIR_REFERENCE_ROOT", "/root/reference")
_DATA = ["World_Age_2019.csv", "AgeSpecificFertility.csv", "Contact_Matrices",
         "c_age.txt", "c_diabetes.txt", "c_hypertension.txt",
         "comorbidity_age_intervals.txt", "validation_data", "Houseshold.csv"]

_scratch = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "seir_individual.py"))


def scratch_dir():
    """Scratch CWD with symlinks to the reference's data files."""
    global _scratch
    if _scratch is None:
        _scratch = tempfile.mkdtemp(prefix="seir_ref_cwd_")
        for name in _DATA:
            src = os.path.join(REFERENCE_ROOT, name)
            if os.path.exists(src):
                os.symlink(src, os.path.join(_scratch, name))
    return _scratch


def load():
    """Import the reference's `seir_individual` module (numba-compiled)."""
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    os.chdir(scratch_dir())
    import seir_individual as ref_seir  # noqa: the reference module
    assert os.path.realpath(ref_seir.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), \
        "picked up a seir_individual that is not the reference: %s" % ref_seir.__file__
    return ref_seir


def driver_namespace(n, country="Italy", stop_line=414):
    """Exec lines 1..stop_line of the reference's run_simulation.py (parameter
    construction only; the driver itself would run 30 x 10M agents at import)
    with params['n'] overridden.  Returns the resulting namespace."""
    import numpy as np
    load()
    src = open(os.path.join(REFERENCE_ROOT, "run_simulation.py")).read().split("\n")[:stop_line]
    src = "\n".join(src)
    src = src.replace("params['n'] = 10000000.", "params['n'] = %d." % n)
    src = src.replace("country = 'Italy'", "country = %r" % country)
    # the driver imports run_complete_simulation from the reference; fine.
    ns = {"__name__": "ref_run_simulation_params"}
    np.random.seed(0)
    exec(compile(src, "run_simulation.py[1:%d]" % stop_line, "exec"), ns)
    return ns
