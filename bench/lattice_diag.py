"""Diagnostic: the tie-heavy lattice golden case, entry point by entry point (which differ, where)."""
import numpy as np
from grid_interpolation_b200 import interpolation as ip
for tag, R in (("f64", np.float64), ("f32", np.float32)):
    z = np.load(f"tests/golden/f90_{tag}.npz")
    def rep(name, got, want):
        bad = np.argwhere(~(np.asarray(got) == np.asarray(want)))
        print(tag, name, "ok" if len(bad) == 0 else f"{len(bad)} differ, first {bad[:6].tolist()}", flush=True)
        for iy, ix in bad[:4]:
            print("    target", float(z["lat_x_axis"][ix]), float(z["lat_y_axis"][iy]), "got", got[iy, ix], "want", want[iy, ix])
    for order in ("F", "C"):
        rep("bary " + order, ip.barycentric_interpolation(z["lat_points"], z["lat_data"], z["lat_x_axis"], z["lat_y_axis"],
            z["lat_simplices"], z["lat_neighbours"], order=order, dtype=R), z["lat_bary_field"])
    for name in ("lat", "dup"):
        P, D = z[name + "_points"], z[name + "_data"]
        for k in (1, 4, 6):
            rep(f"{name} kd k{k}", ip.idw_kdtree(P.T, D, z["lat_x_axis"], z["lat_y_axis"], k, dtype=R), z[f"{name}_kd_field_k{k}"])
            rep(f"{name} kd-refvisits k{k}", ip.idw_kdtree_ex(P.T, D, z["lat_x_axis"], z["lat_y_axis"], k, reference_visits=True, dtype=R), z[f"{name}_kd_field_k{k}"])
            rep(f"{name} esrg k{k}", ip.idw_esrg_nearest(P, D, z["lat_x_axis"], z["lat_y_axis"], 1.0, k, dtype=R), z[f"{name}_esrg_field_k{k}"])
            rep(f"{name} esrg-reforder k{k}", ip.idw_esrg_nearest_ex(P, D, z["lat_x_axis"], z["lat_y_axis"], 1.0, k, reference_order=True, dtype=R), z[f"{name}_esrg_field_k{k}"])
