#!/usr/bin/env python
"""Benchmark of the HiFiLES per-RK-stage residual hot path on B200 (BASELINE.json config 3).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 64] [--order 4] [--impl ours|reference]

One "step" is one full time step of the low-storage SSP-RK34 scheme = 4 passes of the hot path (CalcResidual +
AdvanceSolution) over the whole mesh; the metric counts scalar DOF-RK-stage updates per second over all GPUs.
Workload at N = 1: 3-D Taylor-Green vortex, Re 1600, 64^3 linear hexahedra, P = 4, HLLC + LDG, periodic (synthetic mesh
from hifiles-solver_b200/meshgen.py, analytic initial condition: no dataset involved).  N > 1: the same 64^3 mesh split
into N bricks (strong scaling), halo exchange over NCCL.
See DESIGN.md "Measurement" for the definition of every key of the JSON line."""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_DOF_STAGE = {  # algorithmic bytes per DOF-RK-stage (DESIGN.md "Measurement")
    # a_RK words of state traffic + (Nf/Nu) * (2 + 2*n_dims) face words, 8 bytes each
    "stage": lambda a_rk, nf_nu, nd, visc: 8.0 * (a_rk + nf_nu * (2 + (2 * nd if visc else 0))),
    # share of the dominant kernel k_resid: state traffic + write own face u + read neighbour face u + read neighbour grad
    "k_resid": lambda a_rk, nf_nu, nd, visc: 8.0 * (a_rk + nf_nu * (2 + (nd if visc else 0))),
}
A_RK = {0: 2.0, 1: 2.5, 2: 2.5, 3: 4.0, 4: 4.0}


def load_package():
    import conftest
    return conftest.load_package()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_case(workdir, n, order, blocks=(1, 1, 1), **over):
    hb = load_package()
    import importlib
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    mesh = os.path.join(workdir, "tgv_%d.neu" % n)
    if not os.path.exists(mesh):
        mg.hex_box(mesh, n)
    # dt: the shipped input's 1.440389e-5 (15^3, P = 1) scaled to stay CFL-stable
    dt = 1.440389e-5 * (15.0 / n) * (3.0 / (2 * order + 1))
    opts = dict(order=order, adv_type=2, dt=dt, riemann_solve_type=3, viscous=1)
    opts.update(over)
    inp = mg.write_input(os.path.join(workdir, "input_tgv_%d_p%d" % (n, order)), os.path.basename(mesh), **opts)
    return hb, mg, inp


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "HiFiLES_ref")


def _ref_time_comp(workdir, inp, env, warmup=0):
    """One run of the unmodified reference binary; returns the seconds per time step between its monitored steps, read from the
    last column of history.plt (Time_Comp: the reference's own clock() since start, in minutes; reference src/output.cpp:2403-2406,
    src/HiFiLES.cpp:334-335).  Differences of rows exclude the set-up; the first `warmup` steps after the first row are left out."""
    r = subprocess.run([REF_BIN, os.path.basename(inp)], cwd=workdir, env=env, capture_output=True, text=True)
    hist = os.path.join(workdir, "history.plt")
    if r.returncode != 0 or not os.path.exists(hist):
        return None
    rows = [l for l in open(hist).read().splitlines() if l and l[0].isdigit()]
    if len(rows) < 2 + warmup:
        return None
    t = [float(l.split(",")[-1]) * 60.0 for l in rows]
    return (t[-1] - t[warmup]) / (len(rows) - 1 - warmup)


def cpu_reference_rate(order, n_ref, replicas=1, steps=1, warmup=0):
    """Times the UNMODIFIED reference CPU solver (oracle/_ref/HiFiLES_ref, built from /root/reference by oracle/build_ref.sh) on the
    bounded sample SURVEY.md section 8(d) names: the Taylor-Green case on n_ref^3 hexahedra (15^3 = the size of the mesh the reference
    ships), same order and options as the GPU workload, two time steps; seconds per step = difference of the reference's own
    Time_Comp between the two monitored steps.  replicas > 1: that many concurrent serial runs, one per host core, each on its own
    domain -- the reference has no threading and its MPI build needs MPI + ParMETIS (absent), so concurrent serial domains are its
    multi-core run without the halo exchange, an upper bound for it on these cores; rate = all replicas' updates / slowest replica."""
    import util
    from concurrent.futures import ThreadPoolExecutor
    if not (os.path.exists(REF_BIN) and util.have_reference()):
        return None
    work = tempfile.mkdtemp(prefix="hf_cpu_")
    try:
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR, OMP_NUM_THREADS="1")
        dirs = []
        for r in range(replicas):
            d = os.path.join(work, "r%d" % r)
            os.makedirs(d)
            _, _, inp = make_case(d, n_ref, order, n_steps=1 + warmup + steps, monitor_res_freq=1, plot_freq=1000000, restart_dump_freq=1000000)
            dirs.append((d, inp))
        with ThreadPoolExecutor(max_workers=replicas) as ex:
            secs = list(ex.map(lambda a: _ref_time_comp(a[0], a[1], env, warmup), dirs))
        if any(x is None for x in secs):
            return None
        dof = n_ref ** 3 * (order + 1) ** 3 * 5
        sec = max(max(secs), 1e-9)
        n_rk = 4
        what = ("TGV %d^3 hex P=%d, HLLC + LDG, SSP-RK34: seconds per time step (4 RK stages) over %d step(s) after %d warm-up step(s), from the rows of the "
                "reference's own Time_Comp" % (n_ref, order, steps, warmup))
        sample = ("unmodified reference, serial: " + what) if replicas == 1 else (
            "%d concurrent serial runs (one per host core) of the unmodified reference, each " % replicas + what +
            "; no halo exchange, so an upper bound for the reference's MPI build on these cores")
        return dict(value=replicas * dof * n_rk / sec / 1e9, seconds=sec, dof=dof * replicas, sample=sample)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def host_cores():
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        n = os.cpu_count()
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return max(1, min(int(n or 1), 64))


def workload_text(n, order, n_rk=4):
    return ("3-D Taylor-Green vortex Re=1600, %d^3 hexahedra (global), P=%d, HLLC + LDG(beta=0.5), SSP-RK34, periodic; "
            "1 step = 1 time step = %d RK stages" % (n, order, n_rk))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps K --warmup W as for the GPU arm, each step one time step of the bounded sample (the shipped mesh size, 15^3).  A serial
    # calibration run (1 step) sizes them: W + K steps must fit about two and a half minutes on this host, else both shrink in proportion (the
    # line prints the numbers actually used)
    serial = cpu_reference_rate(args.order, args.cpu_n)
    cores = host_cores()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if serial is not None:
        fit = max(1, int(150.0 / max(serial["seconds"] * 1.3, 1e-3)))  # concurrent replicas run a little slower than one
        if steps + warmup > fit:
            warmup = max(0, min(warmup, fit // 5))
            steps = max(1, fit - warmup)
    res = cpu_reference_rate(args.order, args.cpu_n, cores, steps, warmup) if serial is not None else None
    if res is None:
        res, cores, steps, warmup = serial, 1, 1, 0
    n = args.n
    cfg = {"workload": workload_text(n, args.order), "sample": None}
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present"}))
        return
    cfg["sample"] = res["sample"]
    line = {"impl": "reference", "metric": "GDOF-RK-stage updates/s (TGV hex P=%d)" % args.order, "value": res["value"], "unit": "GDOF-stage/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": res["value"], "unit": "GDOF-stage/s", "cores": cores, "kind": "reference", "sample": res["sample"],
                             "serial_value": serial["value"]},
            "e2e": {"value": res["value"], "unit": "GDOF-stage/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ---- the other BASELINE.json configurations (element types without fused kernels: the staged kernels run) ------------------------------
EULER_VORTEX = dict(ic_form=0, test_case=1, dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, u_c_ic=1., v_c_ic=1., w_c_ic=0., p_c_ic=1., rho_c_ic=1.)
TWO_PI = 6.2831853071795862
OTHER_CONFIGS = {
    # number: (description, mesh generator, default size, CPU sample size, mesh kwargs, input options)
    1: ("2-D Euler isentropic vortex, %s^2 quadrilaterals, P=3, Rusanov, RK45, periodic (BASELINE config 1)", "quad_box", 512, 192, {},
        dict(order=3, adv_type=3, riemann_solve_type=0, viscous=0, dt=1e-4, **EULER_VORTEX)),
    2: ("2-D Navier-Stokes on a mixed triangle / quadrilateral mesh (%s^2 cells, half of them split), P=3, Rusanov + LDG, SSP-RK34, periodic "
        "(BASELINE config 2's discretisation on a synthetic periodic mesh)", "mixed_box_2d", 384, 128, dict(kind="mixed", lengths=(TWO_PI, TWO_PI), origin=(0., 0.)),
        dict(order=3, adv_type=2, riemann_solve_type=0, viscous=1, dt=1e-6, dz_cyclic=None)),
    4: ("3-D Navier-Stokes on a mixed prism / tetrahedron mesh (%s^3 cells), P=3, RoeM + LDG with over-integration (polynomial de-aliasing), SSP-RK34, periodic "
        "(BASELINE config 4's discretisation on a synthetic periodic mesh)", "mixed_box_3d", 24, 10, dict(kind="pritet"),
        dict(order=3, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-6, over_int=1, over_int_order=5)),
    5: ("supersonic wall-bounded LES on a mixed hexahedron / prism mesh (%s^3 cells): Mach 1.8 inflow (sup_in), sup_out outflow, characteristic far field, adiabatic "
        "wall with the Werner-Wengle wall model, WALE sub-grid model, Persson sensor + exponential filter after every stage, P=3, HLLC + LDG, SSP-RK34 "
        "(BASELINE config 5's discretisation on a synthetic mesh)", "mixed_box_3d", 24, 8,
        dict(kind="hexpri", lengths=(1.5, 1., 2.), bcs={"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Far"}),
        dict(order=3, adv_type=2, riemann_solve_type=3, viscous=1, ic_form=1, dt=1e-8, fix_vis=0, Mach_c_ic=1.8, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0.02, T_c_ic=290., rho_c_ic=1.2,
             Mach_free_stream=1.8, rho_free_stream=1.2, T_free_stream=290., L_free_stream=1., dx_cyclic=None, dy_cyclic=1., dz_cyclic=None, bc_In_type="sup_in",
             bc_In_p_static=101000., bc_In_mach=1.8, bc_In_T_static=290., bc_In_nx=1., bc_In_ny=0., bc_In_nz=0., bc_Out_type="sup_out", bc_Wall_type="adiabat_wall",
             bc_Wall_use_wm=1, bc_Far_type="char", bc_Far_p_static=101000., bc_Far_mach=1.8, bc_Far_T_static=290., bc_Far_nx=1., bc_Far_ny=0., bc_Far_nz=0., LES=1,
             SGS_model=1, C_s=0.325, filter_ratio=2.0, wall_model=1, shock_cap=1, s0=1e-9, expf_cutoff=1, calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
    # not a BASELINE configuration: the headline discretisation (config 3) on a mesh WITH boundary faces -- the sum-factorised generation-9
    # kernels with the ghost states evaluated in the face kernel
    6: ("wall-bounded 3-D Navier-Stokes on %s^3 hexahedra (periodic in x, y; adiabatic wall below, characteristic outflow above), P=4, HLLC + LDG, SSP-RK34 "
        "(BASELINE config 3's discretisation with boundary faces)", "hex_box", 48, 10,
        dict(lengths=(1., 1., 2.), bcs={"x-": "Cyclic", "x+": "Cyclic", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Far"}),
        dict(order=4, adv_type=2, riemann_solve_type=3, viscous=1, ic_form=1, dt=1e-7, Mach_c_ic=0.2, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17,
             Mach_free_stream=0.2, rho_free_stream=1.17, T_free_stream=300., L_free_stream=1., dx_cyclic=1., dy_cyclic=1., dz_cyclic=None, bc_Wall_type="adiabat_wall",
             bc_Far_type="sub_out_char", bc_Far_p_static=100500.)),
}


def make_other_case(workdir, config, n, **over):
    hb = load_package()
    import importlib
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    desc, gen, _, _, mkw, opts = OTHER_CONFIGS[config]
    mesh = os.path.join(workdir, "cfg%d_%d.neu" % (config, n))
    if not os.path.exists(mesh):
        getattr(mg, gen)(mesh, n, **mkw)
    o = dict(opts)
    o.update(over)
    inp = mg.write_input(os.path.join(workdir, "input_cfg%d_%d" % (config, n)), os.path.basename(mesh), **o)
    return hb, mg, inp, desc % n


def run_other_config(args):
    """Configurations 1, 2, 4, 5 on one GPU: same metric, same JSON line; the workload runs in fast mode (blocked element kernels around
    the interface kernels), timed with the state resident in HBM, then end to end with host buffers; CPU baseline = the unmodified
    reference on a bounded sample of the same configuration."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--config 1/2/4 run on one GPU")
    cfg = args.config
    n = args.n if args.n_given else OTHER_CONFIGS[cfg][2]
    work = tempfile.mkdtemp(prefix="hf_bench_cfg_")
    hb, mg, inp, desc = make_other_case(work, cfg, n)
    t_setup = time.time()
    run = hb.Run(inp)
    t_setup = time.time() - t_setup
    n_rk = int(run.scalar("n_rk"))
    adv_type = int(run.scalar("adv_type"))
    n_dims = int(run.scalar("n_dims"))
    visc = int(run.scalar("viscous")) != 0
    types = run.ele_types()
    shapes = {t: run.download(t, "disu_upts").shape for t in types}
    dof = float(sum(np.prod(shapes[t]) for t in types))
    fpts = {"tri": lambda p: 3 * (p + 1), "quad": lambda p: 4 * (p + 1), "tet": lambda p: 2 * (p + 1) * (p + 2), "pri": lambda p: (p + 1) * (p + 2) + 3 * (p + 1) ** 2,
            "hex": lambda p: 6 * (p + 1) ** 2}
    # algorithmic bytes per DOF-stage of SURVEY.md 8(d), face term weighted over the element types by their DOF
    bpd = sum(np.prod(shapes[t]) * BYTES_PER_DOF_STAGE["stage"](A_RK[adv_type], fpts[t](args.order_cfg) / shapes[t][0], n_dims, visc) for t in types) / dof

    def sync():
        run.sync()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        run.run(1, fused=True)
    sync()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = run.launch_count()
    run.timer_start()
    run.run(args.steps, fused=True)
    ms = run.timer_stop()
    sync()
    launches = run.launch_count() - l0
    clocks = sampler.stop()
    value = dof * n_rk * args.steps / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    ach = value * bpd
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": None, "algorithmic_bytes_per_dof_stage": bpd, "peak_source": peak_src,
                "note": "whole stage (element kernels + interface kernels) against the algorithmic bytes of SURVEY 8(d), face term weighted over the element types"}
    # dense-operator element types (triangles, tetrahedra, prisms): SURVEY 8(d) names the FP64 pipe as the relevant bound.  Algorithmic
    # flops of the dense operator chain per element and stage: 2 F [Nf Nu (opp_0) + D Nf Nu (opp_1) + D Nu^2 (opp_2) + Nu Nf (opp_3)
    # + viscous: D Nu^2 (opp_4) + D Nu Nf (opp_5) + D Nf Nu (opp_6)]; peak = the DMMA rate measured on this pool's B200
    # (profiles/microbench/dmma_b200.txt: 62.9 FMA/clk/SM = 36.6 TFLOP/s, which is also the DFMA rate)
    flops = 0.0
    for t in types:
        nu_t, ne_t, nf_t = shapes[t]
        nfp_t = fpts[t](args.order_cfg)
        per_ele = 2.0 * nf_t * (nfp_t * nu_t * (1 + n_dims) + n_dims * nu_t ** 2 + nu_t * nfp_t + (n_dims * nu_t ** 2 + 2 * n_dims * nu_t * nfp_t if visc else 0))
        flops += per_ele * ne_t
    tf = flops * n_rk * args.steps / (ms * 1e-3) / 1e12
    roofline["fp64"] = {"achieved": tf, "peak": 36.6, "unit": "TFLOP/s", "frac": tf / 36.6, "algorithmic_flop_per_dof_stage": flops / dof,
                        "peak_source": "profiles/microbench/dmma_b200.txt (measured DMMA m8n8k4 rate, B200)"}
    if tf / 36.6 > ach / peak:
        roofline["bound"] = "tensor"
        roofline["note"] += "; dense operators: the FP64 tensor-core view (roofline.fp64) is the tighter bound"
    e2e = None
    if not args.no_e2e:
        lib = hb.lib()
        ids = {"tri": 0, "quad": 1, "tet": 2, "pri": 3, "hex": 4}
        host = {t: torch.empty(int(np.prod(shapes[t])), dtype=torch.float64, pin_memory=True) for t in types}
        ck = lambda st: (_ for _ in ()).throw(RuntimeError(lib.hf_dev_last_error().decode())) if st != 0 else None
        for t in types:
            ck(lib.hf_dev_download(run.ctx, ids[t], 0, ctypes.c_void_p(host[t].data_ptr()), host[t].numel()))
        e2e_steps = max(2, min(args.steps, 10))
        nbytes = int(dof) * 8

        def timed(mode):
            sync()
            t0 = time.perf_counter()
            if mode == "overlapped":
                for t in types:
                    ck(lib.hf_dev_upload_begin(run.ctx, ids[t], 0, ctypes.c_void_p(host[t].data_ptr()), host[t].numel()))
                for i in range(e2e_steps):
                    for t in types:
                        ck(lib.hf_dev_upload_commit(run.ctx, ids[t]))
                    if i + 1 < e2e_steps:
                        for t in types:
                            ck(lib.hf_dev_upload_begin(run.ctx, ids[t], 0, ctypes.c_void_p(host[t].data_ptr()), host[t].numel()))
                    run.run(1, fused=True)
                    run.norm_residual()
            else:
                for _ in range(e2e_steps):
                    for t in types:
                        ck(lib.hf_dev_upload(run.ctx, ids[t], 0, ctypes.c_void_p(host[t].data_ptr()), host[t].numel()))
                    run.run(1, fused=True)
                    for t in types:
                        ck(lib.hf_dev_download(run.ctx, ids[t], 0, ctypes.c_void_p(host[t].data_ptr()), host[t].numel()))
            sync()
            return dof * n_rk * e2e_steps / (time.perf_counter() - t0) / 1e9

        timed("overlapped")
        v_metric = timed("overlapped")
        v_full = timed("roundtrip")
        e2e = {"value": v_metric, "unit": "GDOF-stage/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 8 * int(run.scalar("n_dims") + 2), "steps": e2e_steps,
               "note": "per step, inside the timed region: the step's input solution host->device from pinned memory (two-phase upload: the copy of step i+1 "
                       "overlaps the stages of step i), all RK stages, residual norm device->host",
               "full_state_roundtrip": {"value": v_full, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes}}
    finite = bool(np.all(np.isfinite(run.norm_residual())))
    if run.fused_status() == "available":
        variant = run.fused_variant()
    elif run.elem_status() == "available" and not os.environ.get("HF_NO_ELEM"):
        variant = "blocked element kernels (k_elem_grad + k_elem_resid per element type, tensor-core operator products) + staged interface kernels"
    else:
        variant = "staged (%s; %s)" % (run.fused_status(), run.elem_status())
    run.close()
    roofline["kernel"] = "all kernels of one RK stage: " + variant
    cpu = None
    if not args.no_cpu:
        import util
        n_cpu = OTHER_CONFIGS[cfg][3]
        d = os.path.join(work, "cpu")
        os.makedirs(d)
        _, _, cinp, cdesc = make_other_case(d, cfg, n_cpu, n_steps=2, monitor_res_freq=1)
        sec = _ref_time_comp(d, cinp, dict(os.environ, HIFILES_HOME=util.REF_DIR)) if os.path.exists(REF_BIN) else None
        if sec:
            with hb.Run(cinp, host_only=True) as h:
                cdof = float(sum(np.prod(h.host_array(t + ".disu_upts").shape) for t in h.ele_types()))
            cpu = {"value": cdof * n_rk / sec / 1e9, "unit": "GDOF-stage/s", "cores": 1, "kind": "reference",
                   "sample": "unmodified reference, serial: " + cdesc + "; one time step between two rows of its own Time_Comp"}
    line = {"metric": "GDOF-RK-stage updates/s (BASELINE config %d)" % cfg if cfg != 6 else "GDOF-RK-stage updates/s (config 3's discretisation, hexahedra with boundary faces)", "value": value, "unit": "GDOF-stage/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + "; 1 step = 1 time step = %d RK stages" % n_rk, "kernels": variant, "elements": {t: int(shapes[t][1]) for t in types}, "dof_total": dof,
                       "l2": "no flush needed: solution %.2f GB plus the staged intermediates >> 126 MB L2" % (dof * 8 / 1e9), "setup_s": round(t_setup, 1)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "residual_finite": finite}
    print(json.dumps(line))
    shutil.rmtree(work, ignore_errors=True)


def parity_gate(hb, mg, dist, rank, world, work, order):
    """Before anything is timed: a small Taylor-Green case (8^3 elements on several GPUs, 6^3 on one; same order, fluxes and RK scheme as
    the workload; 3 time steps) runs through the SAME kernels and halo exchange as the timed run -- on N > 1 GPUs once per partition
    kind (bricks, METIS k-way) -- and is compared on rank 0 with the single-domain run of the fused kernels and with the unmodified
    reference CPU solver (oracle/_ref/ref_dump).  Returns the dict that goes into the JSON line as "parity"."""
    import numpy as np
    import torch
    import util
    n, steps = (8 if world > 1 else 6), 3
    d = os.path.join(work, "parity")
    if rank == 0:
        os.makedirs(d, exist_ok=True)
    if dist is not None:
        dist.barrier()
    inp = None
    if rank == 0:
        _, _, inp = make_case(d, n, order)
    if dist is not None:
        dist.barrier()
        if rank != 0:
            _, _, inp = make_case(d, n, order)
    nu = (order + 1) ** 3

    def scaled_err(got, ref):
        sc = np.abs(ref).reshape(-1, 5).max(0)
        sc[1:4] = sc[1:4].max()
        return float((np.abs(got - ref).reshape(-1, 5).max(0) / sc).max())

    def gathered(part_kind):
        part, nccl_id = None, None
        if world > 1:
            part = mg.block_partition(n, mg.blocks_for(world)) if part_kind == "bricks" else None
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.tensor(list(hb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(idt, src=0)
            nccl_id = bytes(idt.cpu().tolist())
        with hb.Run(inp, rank=rank, nproc=world, part=part, nccl_id=nccl_id) as run:
            variant = run.fused_variant() if run.fused_status() == "available" else "staged: " + run.fused_status()
            run.run(steps, fused=True)
            u = run.download("hex", "disu_upts")
            gid = run.host_array("hex.ele2global_ele")
            res = run.norm_residual()
        full = torch.zeros((n ** 3, nu, 5), dtype=torch.float64, device="cuda")
        full[torch.from_numpy(gid.astype(np.int64)).cuda()] = torch.from_numpy(np.ascontiguousarray(u.transpose(1, 0, 2))).cuda()
        if dist is not None:
            dist.all_reduce(full)
        return full.cpu().numpy(), variant, res

    out = {"case": "TGV %d^3 hex P=%d, HLLC + LDG, SSP-RK34, %d time steps" % (n, order, steps), "measure": "max over fields of max|a-b| / max|b| (momentum components share a scale)"}
    runs = {}
    for kind in (("bricks", "metis") if world > 1 else ("none",)):
        runs[kind] = gathered(kind)
    if rank == 0:
        single, ref_u, ref_res = None, None, None
        if world > 1:
            with hb.Run(inp) as one:
                one.run(steps, fused=True)
                us, gs = one.download("hex", "disu_upts"), one.host_array("hex.ele2global_ele")
            single = np.zeros((n ** 3, nu, 5))
            single[gs] = us.transpose(1, 0, 2)
        if util.have_reference():
            r = util.run_reference(inp, steps, stagewise=False)
            ref_u = np.ascontiguousarray(r["final.hex.disu_upts"].transpose(1, 0, 2))  # the serial reference numbers its elements globally
            ref_res = r["history.norm_residual"][:, -1]
        for kind, (got, variant, res) in runs.items():
            e = {"kernels": variant}
            if single is not None:
                e["vs_single_domain"] = scaled_err(got, single)
            if ref_u is not None:
                e["vs_reference"] = scaled_err(got, ref_u)
                e["residual_norm_vs_reference"] = float(np.abs(np.asarray(res) - ref_res).max() / np.abs(ref_res).max())
            out["partition " + kind if world > 1 else "single domain"] = e
        worst = max([v.get("vs_reference", 0.) for v in out.values() if isinstance(v, dict)] + [v.get("vs_single_domain", 0.) for v in out.values() if isinstance(v, dict)])
        out["max"] = worst
        out["ok"] = bool(worst <= 1e-12)
        out["reference"] = "oracle/_ref/ref_dump (unmodified reference CPU solver)" if ref_u is not None else "absent"
    if dist is not None:
        dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", dest="n", type=int, default=None, help="elements per direction of the global mesh (default 64 for config 3)")
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5, 6], help="BASELINE.json configuration: 3 (default) = TGV hex P=4, the headline; 1, 2, 4, 5 = the "
                    "quad / mixed 2-D / mixed 3-D / wall-bounded LES configurations through the blocked element kernels, one GPU; 6 = config 3's discretisation on a "
                    "hexahedral mesh with walls (sum-factorised kernels with boundary faces)")
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-n", type=int, default=15, help="elements per direction of the CPU baseline sample (15 = the reference's shipped TGV mesh)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (small partitioned case against single domain and reference)")
    ap.add_argument("--staged", action="store_true", help="time the staged (reference-order) kernels instead of the fused ones")
    ap.add_argument("--partition", default="bricks", choices=["bricks", "metis"], help="N > 1: brick partition of the cube, or METIS k-way of the dual graph (the reference's ParMETIS call)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.n_given = args.n is not None
    if args.n is None:
        args.n = 64
    if args.config != 3:
        if args.impl == "reference":
            raise SystemExit("--impl reference times configuration 3")
        args.order_cfg = 4 if args.config == 6 else 3
        return run_other_config(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    work = os.environ.get("HF_BENCH_DIR") or tempfile.mkdtemp(prefix="hf_bench_")
    if world > 1:
        # all ranks must read the same files: rank 0 creates the directory, the others receive its name
        obj = [work if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        work = obj[0]
    hb = mg = inp = None
    if rank == 0:
        hb, mg, inp = make_case(work, args.n, args.order)
    if world > 1:
        dist.barrier()
        if rank != 0:
            hb, mg, inp = make_case(work, args.n, args.order)
    part, nccl_id = None, None
    if world > 1:
        part = mg.block_partition(args.n, mg.blocks_for(world)) if args.partition == "bricks" else None
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(hb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, src=0)
        nccl_id = bytes(idt.cpu().tolist())

    parity = None
    if not args.no_parity:
        parity = parity_gate(hb, mg, dist, rank, world, work, args.order)

    t_setup = time.time()
    run = hb.Run(inp, rank=rank, nproc=world, part=part, nccl_id=nccl_id)
    t_setup = time.time() - t_setup
    if args.staged:
        run.set_mode(False)
    fused = (not args.staged) and run.fused_status() == "available"
    n_eles = run.n_eles("hex")
    nu = (args.order + 1) ** 3
    dof_local = n_eles * nu * 5
    n_rk = int(run.scalar("n_rk"))
    adv_type = int(run.scalar("adv_type"))

    def barrier():
        run.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        run.run(1, fused=True)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = run.launch_count()
    barrier()
    run.timer_start()
    run.run(args.steps, fused=True)
    ms = run.timer_stop()
    barrier()
    launches = run.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        d = torch.tensor([float(dof_local)], dtype=torch.float64, device="cuda")
        dist.all_reduce(d)
        dof_total = float(d.item())
    else:
        dof_total = float(dof_local)
    value = dof_total * n_rk * args.steps / (ms * 1e-3) / 1e9

    # ---- dominant kernel, timed launch by launch with CUDA events on its stream ------------------------------------------
    roofline = None
    peak, peak_src = measured_peak()
    if fused:
        run.kernel_timer(True)
        ksteps = max(2, args.steps // 2)
        run.run(ksteps, fused=True)
        kms, kn = run.kernel_timer(False)
        # one k_resid pass over all elements of the rank per RK stage (two launches when the rank has partition faces:
        # interior elements, then partition-adjacent ones); the time below is the sum of the launches of one stage
        per_launch_s = kms * 1e-3 / (ksteps * n_rk)
        bpd = BYTES_PER_DOF_STAGE["k_resid"](A_RK[adv_type], 6.0 / (args.order + 1), 3, True)
        ach = dof_local * bpd / per_launch_s / 1e9
        stage_b = BYTES_PER_DOF_STAGE["stage"](A_RK[adv_type], 6.0 / (args.order + 1), 3, True)
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "kernel": "k_resid (fused residual + RK update + next-stage face values)", "launch_ms": per_launch_s * 1e3,
                    "algorithmic_bytes_per_dof_stage": bpd, "peak_source": peak_src,
                    "whole_stage": {"algorithmic_bytes_per_dof_stage": stage_b,
                                    "achieved": (dof_total / world) * n_rk * args.steps / (ms * 1e-3) * stage_b / 1e9,
                                    "frac": (dof_total / world) * n_rk * args.steps / (ms * 1e-3) * stage_b / 1e9 / peak}}
        # DRAM bytes of one k_resid launch from the ncu capture of the same command (profiles/), per element x this rank's elements
        tr = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if os.path.exists(tr):
            try:
                roofline["traffic"] = json.load(open(tr))["k_resid_dram_bytes_per_element"] * n_eles
            except Exception:
                pass

    # ---- end to end through the reference-facing API with HOST buffers -----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        lib = hb.lib()
        shape = (nu, n_eles, 5)
        nbytes = int(np.prod(shape)) * 8
        host = torch.empty(int(np.prod(shape)), dtype=torch.float64, pin_memory=True)
        ctx = run.ctx
        ck = lambda st: (_ for _ in ()).throw(RuntimeError(lib.hf_dev_last_error().decode())) if st != 0 else None
        ck(lib.hf_dev_download(ctx, 4, 0, ctypes.c_void_p(host.data_ptr()), host.numel()))
        e2e_steps = max(2, min(args.steps, 10))
        hp, hn = ctypes.c_void_p(host.data_ptr()), host.numel()

        def timed(mode):
            barrier()
            t0 = time.perf_counter()
            if mode == "overlapped":
                # every step's input crosses PCIe inside the timed region; the copy of step i+1 travels on the transfer stream while
                # the stages of step i run (hf_dev_upload_begin / _commit, the two-phase form of eles::cp_disu_upts_cpu_gpu)
                ck(lib.hf_dev_upload_begin(ctx, 4, 0, hp, hn))
                for i in range(e2e_steps):
                    ck(lib.hf_dev_upload_commit(ctx, 4))
                    if i + 1 < e2e_steps:
                        ck(lib.hf_dev_upload_begin(ctx, 4, 0, hp, hn))
                    run.run(1, fused=True)   # CalcResidual + AdvanceSolution x 4
                    run.norm_residual()      # the step's result as the reference reports it: CalcNormResidual (device reduction, 5 doubles to the host)
            else:
                for _ in range(e2e_steps):
                    ck(lib.hf_dev_upload(ctx, 4, 0, hp, hn))  # eles::cp_disu_upts_cpu_gpu: the step's input, synchronous
                    run.run(1, fused=True)
                    if mode == "roundtrip":
                        ck(lib.hf_dev_download(ctx, 4, 0, hp, hn))  # eles::cp_disu_upts_gpu_cpu
                    else:
                        run.norm_residual()
            barrier()
            sec = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([sec], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            return dof_total * n_rk * e2e_steps / sec / 1e9

        timed("overlapped")  # untimed pass: allocates the landing buffer
        v_metric = timed("overlapped")
        v_serial = timed("serial")
        v_full = timed("roundtrip")
        e2e = {"value": v_metric, "unit": "GDOF-stage/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 5 * 8, "steps": e2e_steps,
               "note": "per step, inside the timed region: the step's input solution host->device from pinned memory (two-phase hf_dev_upload_begin/_commit: "
                       "the copy of step i+1 overlaps the stages of step i), 4 RK stages, residual norm device->host (CalcNormResidual)",
               "serial_upload": {"value": v_serial, "note": "same with the synchronous hf_dev_upload (eles::cp_disu_upts_cpu_gpu) in front of every step"},
               "full_state_roundtrip": {"value": v_full, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                                        "note": "synchronous upload, 4 stages, and the whole solution device->host every step (eles::cp_disu_upts_gpu_cpu): the round-1 definition"}}

    res_norm = run.norm_residual()
    finite = bool(np.all(np.isfinite(res_norm)))
    run.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_rate(args.order, args.cpu_n)
        if r is not None:
            cpu = {"value": r["value"], "unit": "GDOF-stage/s", "cores": 1, "kind": "reference", "sample": r["sample"]}

    if rank == 0:
        line = {
            "metric": "GDOF-RK-stage updates/s (TGV hex P=%d)" % args.order, "value": value, "unit": "GDOF-stage/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.n, args.order, n_rk),
                       "kernels": "fused" if fused else "staged", "elements_per_gpu": n_eles, "dof_total": dof_total,
                       "l2": "no flush needed: state per GPU %.2f GB >> 126 MB L2" % (dof_local * 8 / 1e9), "setup_s": round(t_setup, 1),
                       "partition": ("bricks %s" % (mg.blocks_for(world),) if args.partition == "bricks" else "METIS k-way (dual graph)") if world > 1 else "none"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "residual_finite": finite, "parity": parity,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not os.environ.get("HF_BENCH_DIR") and rank == 0:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
