"""Prints the relative error of every compared array for every staged-parity case (no assertions): run on the GPU
box, output kept under profiles/ as the parity record of the round."""
import pathlib
import sys
import tempfile

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).parent))
import conftest  # noqa: E402
import util  # noqa: E402
import test_staged_parity as T  # noqa: E402


def main():
    hb = conftest.load_package()
    import importlib
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    only = sys.argv[1:] or list(T.CASES)
    for name in only:
        tmp = pathlib.Path(tempfile.mkdtemp())
        inp = T.make_case(tmp, mg, name)
        errs = []
        orig = T.check

        def rec(nm, got, ref, tol=T.TOL, scale_by=None):
            err = util.rel_err(got, ref) if scale_by is None else np.abs(got - ref).max() / np.abs(scale_by).max()
            errs.append((nm, err, tol))
            return err
        T.check = rec
        try:
            T.test_methods_one_by_one.__wrapped__ if hasattr(T.test_methods_one_by_one, "__wrapped__") else None
            T.test_methods_one_by_one(tmp, hb, mg, name)
            n1 = len(errs)
            T.test_time_steps_reference_call_sequence(tmp, hb, mg, name)
        finally:
            T.check = orig
        print("== %s" % name)
        for i, (nm, err, tol) in enumerate(errs):
            print("   %-58s %.3e %s" % (nm, err, "" if err <= tol else "  > %.0e" % tol))


if __name__ == "__main__":
    main()
