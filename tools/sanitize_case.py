import sys, os, tempfile, importlib
sys.path.insert(0,'tests')
import conftest
hb = conftest.load_package()
mg = importlib.import_module("hifiles_solver_b200.meshgen")
w = tempfile.mkdtemp()
mg.hex_box(os.path.join(w,"m.neu"), 3)
inp = mg.write_input(os.path.join(w,"input"),"m.neu",order=4,viscous=1,adv_type=2,dt=1e-5,riemann_solve_type=3)
with hb.Run(inp) as run:
    print(run.fused_variant())
    run.run(1, fused=True)
    import numpy as np
    print(np.isfinite(run.download("hex","disu_upts")).all())
