"""Time pm.generate_whitenoise on one GPU (seed-table build on the host + column kernel)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from pmesh_b200.pm import ParticleMesh
    out = {}
    for n in (256, 512, 1024):
        pm = ParticleMesh(BoxSize=1.0, Nmesh=[n, n, n], dtype="f8")
        c = pm.create("complex")
        from pmesh_b200.whitenoise import generate
        generate(c._dev, c.start, c.Nmesh, 1, False)
        pm.ctx.sync()
        t0 = time.perf_counter()
        generate(c._dev, c.start, c.Nmesh, 2, False)
        pm.ctx.sync()
        t1 = time.perf_counter()
        out[n] = {"ms_total": round((t1 - t0) * 1e3, 2), "Gmodes_per_s": round(n * n * (n // 2 + 1) / (t1 - t0) / 1e9, 2)}
        del c, pm
    print(json.dumps({"whitenoise": out}))


if __name__ == "__main__":
    main()
