"""Mesh parameters of the golden fixtures (kept apart from make_golden.py so that CPU tests need no reference)."""
GOLDEN_CASES = {
    "hex3_p2_ns_hllc_rk34": ("hex", 3, {}),
    "hex2_p3_ns_rusanov_rk45": ("hex", 2, {}),
    "hex3_p1_ns_roem_sutherland_rk24": ("hex", 3, {}),
    "quad4_p3_euler_vortex_hllc_rk45": ("quad", 4, {}),
    "quad4_p2_ns_rusanov_euler": ("quad", 4, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.))),
    "tri3_p3_ns_rusanov_rk34": ("tri", 3, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.))),
    "tet1_p2_ns_roem_rk34": ("tet", 1, {}),
    "pri1_p2_ns_hllc_rk45": ("pri", (1, 2, 1), {}),
    # boundary interfaces (tests/golden/make_golden.py)
    "hexbdy_p2_char_suboutsimp_isotherm_adiabat_hllc": ("hex", (3, 2, 3), dict(lengths=(1.5, 1., 1.5), bcs={"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic",
                                                                                                           "z-": "Wall", "z+": "Top"})),
    "hexbdy_p2_supin_supout_slip_isotherm_roem_betaneg": ("hex", (3, 2, 3), dict(lengths=(1.5, 1., 1.5), bcs={"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic",
                                                                                                             "z-": "Wall", "z+": "Top"})),
    "hexbdy_p1_subinchar_suboutchar_adiabat_slipdual": ("hex", (3, 3, 2), dict(lengths=(1.5, 1., 1.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall",
                                                                                                          "z-": "Wall", "z+": "Top"})),
    "quadbdy_p3_euler_subinsimp_suboutsimp_slipdual": ("quad", (6, 5), dict(lengths=(3., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall"})),
    "tribdy_p2_ns_supin_supout_isotherm_adiabat_rusanov": ("tri", (5, 4), dict(lengths=(3., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"})),
}
