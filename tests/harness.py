"""Drives one problem through the CPU oracle or through the CUDA C-ABI with the same call sequence."""
import importlib

import numpy as np

import oracle as O

isph = importlib.import_module("implicit-sph_b200")


def run_oracle(P, F, kind="port", anti=True, singular=O.NULLSPACE, mh=False):
    cs = P["case"]
    o = O.Oracle(P, kinds=cs["kinds"], kernel=cs["kernel"], h_min=cs["h_min"], kind=kind)
    o.set_field(O.F_DENSITY, F["density"]); o.set_field(O.F_VISCOSITY, F["viscosity"]); o.set_field(O.F_PRESSURE, F["pressure"])
    o.set_field(O.F_VSTAR, F["velocity"]); o.set_field(O.F_VELOCITY, F["velocity"]); o.set_field(O.F_FORCE, F["force"])
    o.set_field(O.F_EPS, F["eps"]); o.set_field(O.F_PSI, F["psi"]); o.set_field(O.F_PSI0, F["psi0"])
    o.set_field(O.F_SIGMA, F["sigma"]); o.set_field(O.F_PHI, F["phi"])
    o.compute_pre(normals=cs["has_solid"])
    out = dict(vfrac=o.get_field(O.F_VFRAC), gc=o.get_field(O.F_GC), lc=o.get_field(O.F_LC))
    if cs["has_solid"]:
        out["normal"] = o.get_field(O.F_NORMAL); out["pnd"] = o.get_field(O.F_PND)
    out["rowptr"], out["col"] = o.graph()
    nl, dim = P["nlocal"], P["dim"]
    out["b_poisson"] = o.ns_poisson(cs["dt"], anti=anti, singular=singular, morris_holmes=mh); out["A_poisson"] = o.matrix()
    out["diag_poisson"], out["sld_poisson"] = o.diagonals(); o.invalidate_matrix()
    b0 = np.asfortranarray(F["velocity"][:nl, :dim])
    out["b_helmholtz"] = o.ns_helmholtz(cs["dt"], cs["theta"], b0, anti=anti, morris_holmes=mh); out["A_helmholtz"] = o.matrix(); o.invalidate_matrix()
    out["b_aep"] = o.applied_electric_potential(); out["A_aep"] = o.matrix(); o.invalidate_matrix()      # applied electric potential,
    out["b_solute"] = o.solute_transport(cs["dt"], cs["theta"], 0.7, F["conc"]); out["A_solute"] = o.matrix(); o.invalidate_matrix()   # solute transport
    out["pb_f"] = o.pb_residual(morris_holmes=mh, extra_f=F["pb_extra"][:nl])                     # PB residual: before the Jacobian (no matrix needed)
    out["pb_f_lin"] = o.pb_residual(morris_holmes=mh, linearized=True, gamma=0.1)
    o.pb_jacobian(morris_holmes=mh); out["A_pb"] = o.matrix()
    o.set_field(O.F_PSI, np.cos(P["xw"][:, 0])); o.pb_jacobian(morris_holmes=mh); out["A_pb2"] = o.matrix()
    x = np.random.default_rng(3).standard_normal((nl, 2)); out["spmv_x"] = x; out["spmv_y"] = o.spmv(x)
    dp = np.sin(2 * P["xw"][:nl, 0]) * np.cos(P["xw"][:nl, 1]) + 0.3          # stand-in Poisson solution for the post-solve block
    o.ns_correct(cs["dt"], dp, anti=anti); out["corr_vstar"] = o.get_field(O.F_VSTAR); out["corr_p"] = o.get_field(O.F_PRESSURE); out["corr_dp"] = o.get_field(O.F_DP)
    if cs["has_solid"]:          # boundary-condition row modifiers applied after the Helmholtz functor (pair_isph_corrected.cpp:918-934)
        o.invalidate_matrix(); o.ns_helmholtz(cs["dt"], cs["theta"], b0, anti=anti, morris_holmes=mh); o.boundary_navier_slip(0.7); out["A_slip"] = o.matrix()
        o.invalidate_matrix(); bh = o.ns_helmholtz(cs["dt"], cs["theta"], b0, anti=anti, morris_holmes=mh); out["b_dirichlet"] = o.boundary_dirichlet(bh); out["A_dirichlet"] = o.matrix()
    # advanceTime (v^{n+1} = the corrected vstar left by ns_correct; type 2 is "fixed" where it exists)
    o.set_fixed(FIXED(cs)); o.set_field(O.F_PRESSURE, F["pressure"]); o.set_field(O.F_VELOCITY, F["velocity"])
    o.advance_time(cs["dt"], anti=anti)
    out["adv_dp"] = o.get_field(O.F_DP); out["adv_p"] = o.get_field(O.F_PRESSURE); out["adv_v"] = o.get_field(O.F_VELOCITY); out["adv_x"] = o.get_x()
    o.close()
    return out


def FIXED(cs):
    return [0, 0, 1][:len(cs["kinds"])] if len(cs["kinds"]) >= 3 else [0] * len(cs["kinds"])


def cuda_context(P, F, device=0, device_neighbors=False):
    cs = P["case"]
    c = isph.Context(device)
    c.set_particles(P, kinds=cs["kinds"], kernel=cs["kernel"], h_min=cs["h_min"])
    if device_neighbors:          # the list LAMMPS would hand over is replaced by the one built on the device from the atoms (isph_neighbors_build)
        c.neighbors_build()
    c.field_set(isph.F_DENSITY, F["density"]); c.field_set(isph.F_VISCOSITY, F["viscosity"]); c.field_set(isph.F_PRESSURE, F["pressure"])
    c.field_set(isph.F_VSTAR, F["velocity"]); c.field_set(isph.F_VELOCITY, F["velocity"]); c.field_set(isph.F_FORCE, F["force"])
    c.field_set(isph.F_EPS, F["eps"]); c.field_set(isph.F_PSI, F["psi"]); c.field_set(isph.F_PSI0, F["psi0"])
    c.field_set(isph.F_SIGMA, F["sigma"]); c.field_set(isph.F_PHI, F["phi"])
    return c


def run_cuda(P, F, anti=True, singular=isph.NULLSPACE, mh=False, device=0, device_neighbors=False):
    cs = P["case"]
    c = cuda_context(P, F, device, device_neighbors)
    if device_neighbors:
        noff, neigh = c.neighbors_get()
    c.compute_pre(normals=cs["has_solid"])
    out = dict(vfrac=c.field_get(isph.F_VFRAC), gc=c.field_get(isph.F_GC), lc=c.field_get(isph.F_LC))
    if cs["has_solid"]:
        out["normal"] = c.field_get(isph.F_NORMAL); out["pnd"] = c.field_get(isph.F_PND)
    c.graph_build()
    out["rowptr"], out["col"] = c.graph_get()
    nl, dim = P["nlocal"], P["dim"]
    c.create_load(None, 1)
    c.ns_poisson(cs["dt"], anti=anti, singular=singular, morris_holmes=mh)
    out["b_poisson"] = c.load_get(1)[:, 0]; out["A_poisson"] = c.matrix_get(); out["diag_poisson"], out["sld_poisson"] = c.diagonals_get(); c.matrix_invalidate()
    c.create_load(None, dim); c.load_set(np.asfortranarray(F["velocity"][:nl, :dim]))
    c.ns_helmholtz(cs["dt"], cs["theta"], anti=anti, morris_holmes=mh)
    out["b_helmholtz"] = c.load_get(dim); out["A_helmholtz"] = c.matrix_get(); c.matrix_invalidate()
    c.create_load(None, 1); c.applied_electric_potential(); out["b_aep"] = c.load_get(1)[:, 0]; out["A_aep"] = c.matrix_get(); c.matrix_invalidate()
    c.create_load(None, 1); c.load_set(F["conc"][:nl]); c.solute_transport(cs["dt"], cs["theta"], 0.7)
    out["b_solute"] = c.load_get(1)[:, 0]; out["A_solute"] = c.matrix_get(); c.matrix_invalidate()
    out["pb_f"] = c.pb_residual(morris_holmes=mh, extra_f=F["pb_extra"][:nl])
    out["pb_f_lin"] = c.pb_residual(morris_holmes=mh, linearized=True, gamma=0.1)
    c.pb_jacobian(morris_holmes=mh); out["A_pb"] = c.matrix_get()
    c.field_set(isph.F_PSI, np.cos(P["xw"][:, 0])); c.pb_jacobian(morris_holmes=mh); out["A_pb2"] = c.matrix_get()
    x = np.random.default_rng(3).standard_normal((nl, 2)); out["spmv_x"] = x; out["spmv_y"] = c.matrix_multiply(x)
    dp = np.sin(2 * P["xw"][:nl, 0]) * np.cos(P["xw"][:nl, 1]) + 0.3
    c.ns_correct(cs["dt"], anti=anti, dp=dp); out["corr_vstar"] = c.field_get(isph.F_VSTAR); out["corr_p"] = c.field_get(isph.F_PRESSURE); out["corr_dp"] = c.field_get(isph.F_DP)
    if cs["has_solid"]:
        b0 = np.asfortranarray(F["velocity"][:nl, :dim])
        c.matrix_invalidate(); c.create_load(None, dim); c.load_set(b0); c.ns_helmholtz(cs["dt"], cs["theta"], anti=anti, morris_holmes=mh); c.boundary_navier_slip(0.7); out["A_slip"] = c.matrix_get()
        c.matrix_invalidate(); c.create_load(None, dim); c.load_set(b0); c.ns_helmholtz(cs["dt"], cs["theta"], anti=anti, morris_holmes=mh); c.boundary_dirichlet()
        out["b_dirichlet"] = c.load_get(dim); out["A_dirichlet"] = c.matrix_get()
    c.pair_fixed(FIXED(cs)); c.field_set(isph.F_PRESSURE, F["pressure"]); c.field_set(isph.F_VELOCITY, F["velocity"])
    c.advance_time(cs["dt"], anti=anti)
    out["adv_dp"] = c.field_get(isph.F_DP); out["adv_p"] = c.field_get(isph.F_PRESSURE); out["adv_v"] = c.field_get(isph.F_VELOCITY); out["adv_x"] = c.atoms_get_x()
    out["launches"] = c.launches
    if device_neighbors:
        out["neigh_noff"], out["neigh"] = noff, neigh
    c.close()
    return out
