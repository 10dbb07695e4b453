"""GPU parity of the blocked element kernels (hf_elem.cu): the fast mode of every mesh the sum-factorised hexahedron kernels do not
take -- quadrilaterals, triangles, tetrahedra, prisms, mixed meshes, curved hexahedra, meshes with boundary faces.  Two kernels per
element type and RK stage (k_elem_grad, k_elem_resid: every operator product a batched FP64 tensor-core contraction on a
shared-memory tile) around the staged interface kernels.  Tensor-core accumulation order and fused products differ from the
reference's ascending dgemm sums in the last bits; bar: 1e-12 relative on the solution and the residual norm after three time steps
(BASELINE.json north_star), against the unmodified reference CPU solver (oracle/_ref)."""
import numpy as np
import pytest

import util
from test_staged_parity import CASES, make_case, check

ELEM_CASES = [
    # BASELINE config 1 (quadrilaterals, Euler) and the boundary kinds on quadrilaterals
    "quad_p3_euler_vortex_rk45", "quad_p2_ns_hllc_rk34", "quad_p2_ns_walls_char_out", "quad_p3_euler_slip_supin_supout", "quad_p3_euler_subinsimp_slipdual",
    # BASELINE config 2 (triangles / mixed 2-D meshes, Navier-Stokes)
    "tri_p3_ns_rusanov_rk34", "tri_p2_euler_vortex_hllc_rk45", "mixed_tri_quad_p3_ns_rusanov_walls",
    # BASELINE config 4 (tetrahedra / prisms / hexahedra in one mesh, RoeM, over-integration)
    "tet_p3_ns_roem_rk34", "pri_p3_ns_roem_rk45", "hexpri_p2_ns_roem_rk34", "pritet_p3_ns_roem_rk34",
    "hex_p3_ns_roem_overint", "pritet_p2_ns_roem_overint", "mixed_tri_quad_p3_euler_overint",
    # hexahedra outside the sum-factorised kernels' reach: curved, with boundary faces
    "hex_p2_warped_ns_rk414", "hex_p2_ns_wall_char_periodic", "hex_p2_ns_subinsimp_slipdual",
    # shock capturing after every stage
    "mixed_tri_quad_p3_euler_shockcap", "pritet_p2_ns_hllc_shockcap",
    # CFL time steps (calc_time_step on the device before every step): global on curved triangles with the sensor, local on tetrahedra
    "tri6_p3_ns_curved_supin_wall_cfl_shockcap", "tet_p2_ns_hllc_cfl_local_dt", "quad_p2_ns_cfl_global_dt",
    # LES (BASELINE config 5): the blocked kernels around the staged sub-grid-scale point fluxes -- eddy-viscosity models, filter-based
    # models (calc_sgs_terms at the first stage; SVV replaces the solution), wall models, and config 5's combination with the sensor
    "hex_p3_les_wale_rk34", "pritet_p2_les_wale_rk34", "mixed_tri_quad_p3_les_smagorinsky_walls", "hex_p3_les_wsm_vasilyev",
    "mixed_tri_quad_p3_les_similarity_gaussian", "tet_p2_les_svv_modal", "mixed_tri_quad_p3_les_wale_werner_wengle", "hex_p2_les_wale_loglaw_wall",
    "config5_hexpri_p2_les_wm_shockcap_hllc", "hexpri_p2_les_smagorinsky_periodic",
]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ELEM_CASES)
def test_blocked_element_kernels_vs_reference(tmp_path, hb, meshgen, name, monkeypatch):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    # affine hexahedra with boundary faces take the sum-factorised kernels by default (tests/test_fused_parity.py); here the blocked ones are meant
    monkeypatch.setenv("HF_FUSED_BDY", "0")
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    with hb.Run(inp) as run:
        assert run.fused_status() != "available"
        assert run.elem_status() == "available", run.elem_status()
        n0 = run.launch_count()
        run.run(n_steps, fused=True)
        launches = run.launch_count() - n0
        fast = {t: run.download(t, "disu_upts") for t in run.ele_types()}
        res = run.norm_residual()
    with hb.Run(inp) as run:
        run.set_mode(False)
        n0 = run.launch_count()
        run.run(n_steps, fused=False)
        staged_launches = run.launch_count() - n0
    # the blocked kernels really ran: far fewer launches than one kernel per reference method
    assert launches < 0.6 * staged_launches, (launches, staged_launches)
    # the residual is a derivative of fluxes; behind the sub-grid-scale models' pow(x, 1.25 | 1.5 | 2.5) (device and host libm differ by an
    # ulp there) its norm is held to 1e-11, the solution itself to 1e-12 as everywhere
    check("residual norm", res, ref["history.norm_residual"][:, -1], 1e-11 if CASES[name][3].get("LES") else 1e-12)
    for t in fast:
        check("final disu_upts " + t, fast[t], ref["final." + t + ".disu_upts"], 1e-12)


@pytest.mark.gpu
def test_blocked_kernels_stagewise_residual_and_gradient(tmp_path, hb, meshgen):
    """hf_dev_rk_stage with keep_residual leaves div_tconf_upts of that stage behind (what CalcNormResidual reads): compared with the
    staged kernels' array after the same stage sequence."""
    name = "pritet_p3_ns_roem_rk34"
    inp = make_case(tmp_path, meshgen, name)
    out = {}
    for mode in ("fast", "staged"):
        with hb.Run(inp) as run:
            run.set_mode(mode == "fast")
            for stage in range(4):
                run.rk_stage(stage, 0.0, keep_residual=True)
            out[mode] = {t: (run.download(t, "div_tconf_upts"), run.download(t, "disu_upts")) for t in run.ele_types()}
    for t in out["fast"]:
        check("div_tconf_upts " + t, out["fast"][t][0], out["staged"][t][0], 5e-11)
        check("disu_upts " + t, out["fast"][t][1], out["staged"][t][1], 1e-13)
