"""C1 barycentric per-call time for strip heights of the walk kernel (GI_BARY_SMALL_ROWS probe)."""
import os, sys, time, subprocess
if len(sys.argv) > 1:
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from grid_interpolation_b200 import interpolation as ip
    import workloads
    from oracle import oracle as orc
    for npts in (5000, 100000):
        w = workloads.c1(num_points=npts)
        pf = np.asfortranarray(w.points); sf = np.asfortranarray(w.simplices); nf = np.asfortranarray(w.neighbours)
        want = orc.barycentric_interpolation(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf)
        for order in ("F", "C"):
            got, g = ip.barycentric_interpolation_ex(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf, order=order, debug=True)
            assert np.array_equal(got, want)
            def f(): ip.barycentric_interpolation_ex(pf, w.data_pts, w.x_axis, w.y_axis, sf, nf, order=order)
            for _ in range(5): f()
            t0 = time.perf_counter()
            for _ in range(200): f()
            dt = (time.perf_counter() - t0) / 200 * 1e3
            ip.set_profiling(True); f(); ph = ip.last_timings(); ip.set_profiling(False)
            print(f"rows={sys.argv[1]} pts={npts} grid={w.x_axis.size}x{w.y_axis.size} order={order}: {dt:.3f} ms/call iters_tot={g.iters_tot} "
                  + " ".join(f"{n}={v*1e3:.0f}us" for n, v in ph), flush=True)
else:
    for rows in ("32", "16", "8", "4", "0"):
        env = dict(os.environ)
        if rows == "0":
            env.pop("GI_BARY_SMALL_ROWS", None)   # the library's own choice
        else:
            env["GI_BARY_SMALL_ROWS"] = rows
        subprocess.run([sys.executable, __file__, rows], env=env)
