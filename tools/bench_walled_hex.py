"""Throughput of a hexahedral mesh WITH boundary faces (periodic in x, y; adiabatic wall and characteristic far field in z): the
sum-factorised generation-9 kernels with the ghost states evaluated in the face kernel (default), the blocked element kernels
(HF_FUSED_BDY=0) and the staged kernels.  usage: python tools/bench_walled_hex.py [n] [order] [fast,blocked,staged]"""
import os, sys, tempfile, importlib, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import conftest
hb = conftest.load_package()
mg = importlib.import_module("hifiles_solver_b200.meshgen")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = tempfile.mkdtemp(prefix="hf_walled_")
mg.hex_box(os.path.join(w, "m.neu"), (n, n, n), lengths=(1., 1., 2.), bcs={"x-": "Cyclic", "x+": "Cyclic", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Far"})
inp = mg.write_input(os.path.join(w, "input"), "m.neu", order=order, adv_type=2, riemann_solve_type=3, viscous=1, ic_form=1, dt=1e-7, Mach_c_ic=0.2, nx_c_ic=1., ny_c_ic=0.,
                     nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.2, rho_free_stream=1.17, T_free_stream=300., L_free_stream=1., dx_cyclic=1.,
                     dy_cyclic=1., dz_cyclic=None, bc_Wall_type="adiabat_wall", bc_Far_type="sub_out_char", bc_Far_p_static=100500.)
for mode in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("fast", "blocked", "staged")):
    if mode == "blocked":
        os.environ["HF_FUSED_BDY"] = "0"
    else:
        os.environ.pop("HF_FUSED_BDY", None)
    with hb.Run(inp) as run:
        if mode == "staged":
            run.set_mode(False)
        status = run.fused_status() + " | " + run.elem_status()
        dof = float(np.prod(run.download("hex", "disu_upts").shape))
        fused = mode != "staged"
        run.run(2, fused=fused)
        run.sync()
        run.timer_start()
        run.run(5, fused=fused)
        ms = run.timer_stop()
        ok = bool(np.all(np.isfinite(run.norm_residual())))
    print("hexahedra with walls %d^3 P=%d, HLLC + LDG, RK34, %s kernels: %.3f GDOF-stage/s (%.2f ms per step, residual finite %s) [%s]"
          % (n, order, mode, dof * 4 * 5 / (ms * 1e-3) / 1e9, ms / 5, ok, status), flush=True)
