#!/usr/bin/env python
"""SASS of the NVRTC-compiled pass kernels into profiles/ (no GPU needed: NVRTC + cuobjdump):
   r02_k1_tma.sass   nls_pass (per-launch TMA ring) and nls_pass_persistent, model A*exp(-lam*x)+b, p = 3
   r02_k1b.sass      nls_pass (tiled FP64 DMMA SYRK), sum of 16 Gaussians, p = 48
Encodings are stripped; a mnemonic histogram heads each file."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def dump(tune, rhs, names, out, keep):
    cub = "/tmp/_dump_%s.cubin" % os.path.basename(out)
    env = dict(os.environ, GSLNLS_DUMP_CUBIN=cub)
    if tune:
        env["GSLNLS_TUNE"] = tune
    code = ("import sys; sys.path.insert(0, %r)\nfrom gslnls_b200 import Model\n"
            "Model(%r, %r, ['x'], jac=True, fvv=%r)\n" % (ROOT, rhs, names, len(names) <= 8))
    subprocess.run([sys.executable, "-c", code], env=env, check=True, stderr=subprocess.DEVNULL)
    sass = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
    with open(out, "w") as fh:
        fh.write("# cuobjdump -sass of the NVRTC cubin (sm_100a); GSLNLS_TUNE=%s\n# model: %s\n" % (tune or "default", rhs[:100]))
        for f in keep:
            body = funcs.get(f, [])
            hist = collections.Counter(re.sub(r"^(@!?U?P\d+\s+)?", "", re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", b)).split()[0].rstrip(";")
                                       for b in body if b.strip())
            fh.write("\n# ---- Function : %s  (%d instructions)\n# mnemonic histogram: %s\n" % (
                f, len(body), ", ".join("%s x%d" % kv for kv in hist.most_common(40))))
            fh.write("\n".join(body) + "\n")
    print(out, {f: len(funcs.get(f, [])) for f in keep})


dump("tiled=2,block=416,unroll=3,minb=1,stages=4,fexp=1", bench.FORMULA_RHS, ["A", "lam", "b"],
     os.path.join(ROOT, "profiles", "r02_k1_tma.sass"), ["nls_pass", "nls_pass_persistent"])
rhs, names = bench.gaussmix_formula(16)
dump("", rhs, names, os.path.join(ROOT, "profiles", "r02_k1b.sass"), ["nls_pass"])
