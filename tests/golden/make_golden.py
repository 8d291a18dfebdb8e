#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference tree (run in the build container only).

Nothing here is imported at test time; the GPU box has no /root/reference.  The script
 * parses the 33 formula test problems (data, formula, start, certified target) out of
   /root/reference/R/nls_test.R:169-990 (NIST StRD + Bates/Watts sets),
 * copies the Example-1 data of inst/unit_tests/unit_tests_gslnls.R:256-265,
 * regenerates the README.md Example-2 data (set.seed(1); rnorm) with a restatement of
   R's default RNG (Mersenne-Twister + inversion) that is first checked against the
   Example-1 responses, which come from the same stream,
 * parses the per-iteration traces that README.md prints for Example 2 (lm, lmaccel with
   finite-difference and analytic fvv), and the scalar answers of Examples 1, 3, 4.

Usage:  python tests/golden/make_golden.py  [/root/reference]
"""
import json
import math
import os
import re
import sys

import numpy as np
from scipy.special import ndtri

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# R's default RNG: Mersenne-Twister, set.seed() scrambling, rnorm by inversion
# --------------------------------------------------------------------------------------
class RRandom:
    N, M = 624, 397

    def __init__(self, seed):
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        self.mt = [0] * self.N
        for j in range(self.N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            if j > 0:
                self.mt[j - 1] = s
        self.mti = self.N  # dummy[0] = 624 after FixupSeeds

    def _genrand(self):
        N, M = self.N, self.M
        mt = self.mt
        if self.mti >= N:
            for kk in range(N - M):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            for kk in range(N - M, N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            y = (mt[N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return (y & 0xFFFFFFFF) * 2.3283064365386963e-10

    def unif_rand(self):
        v = self._genrand()
        if v <= 0.0:
            return 0.5 * 2.328306437080797e-10
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return v

    def norm_rand(self):
        BIG = 134217728.0
        u = self.unif_rand()
        u = int(BIG * u) + self.unif_rand()
        return float(ndtri(u / BIG))


# --------------------------------------------------------------------------------------
# tiny R-literal helpers
# --------------------------------------------------------------------------------------
def _match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'"""
    depth = 0
    in_str = None
    for k in range(i, len(s)):
        ch = s[k]
        if in_str:
            if ch == in_str:
                in_str = None
            continue
        if ch in "\"'":
            in_str = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced")


def _num(tok):
    tok = tok.strip()
    tok = re.sub(r"(?<=\d)L$", "", tok)
    return float(eval(tok, {"__builtins__": {}}, {"pi": math.pi, "sqrt": math.sqrt, "exp": math.exp}))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _parse_named_c(body):
    """'a = 1, b = 2e-3' -> (names, values)"""
    names, vals = [], []
    for part in _split_top(body):
        k, v = part.split("=", 1)
        names.append(k.strip())
        vals.append(_num(v))
    return names, vals


def parse_formula_problems(path):
    src = open(path).read()
    lo = src.index('if(identical(name, "Misra1a"))')
    hi = src.index("## function problems") if "## function problems" in src else src.index('identical(name, "Bard")')
    sec = src[lo:hi]
    blocks = re.split(r'identical\(name, "([^"]+)"\)\)\s*\{', sec)
    probs = {}
    for name, body in zip(blocks[1::2], blocks[2::2]):
        m = re.search(r"\.data <- data\.frame\(", body)
        if not m:
            continue
        a = m.end() - 1
        b = _match_paren(body, a)
        cols = {}
        for part in _split_top(body[a + 1:b]):
            k, v = part.split("=", 1)
            v = v.strip()
            assert v.startswith("c("), (name, v[:30])
            inner = v[2:_match_paren(v, 1)]
            cols[k.strip()] = [_num(t) for t in _split_top(inner)]
        fm = re.search(r'\.fn <- as\.formula\("([^"]+)"', body)
        sm = re.search(r"\.start <- c\(", body)
        tm = re.search(r"\.target <- c\(", body)
        if not (fm and sm and tm):
            continue
        sb = body[sm.end():_match_paren(body, sm.end() - 1)]
        tb = body[tm.end():_match_paren(body, tm.end() - 1)]
        pn, sv = _parse_named_c(sb)
        tn, tv = _parse_named_c(tb)
        assert pn == tn, name
        lens = {len(v) for v in cols.values()}
        assert len(lens) == 1, (name, lens)
        probs[name] = {"formula": fm.group(1), "data": cols, "param_names": pn, "start": sv,
                       "target": tv, "n": lens.pop(), "p": len(pn)}
    return probs


def parse_vector(src, varname):
    m = re.search(r"^%s <- c\(" % re.escape(varname), src, re.M)
    b = _match_paren(src, m.end() - 1)
    return [_num(t) for t in _split_top(src[m.end():b])]


def parse_traces(readme):
    """all 'iter k: ssr = s, par = (a, b, c)' runs in README.md, in order of appearance"""
    runs, cur, last = [], [], None
    for line in readme.splitlines():
        m = re.match(r"#> iter\s+(\d+): ssr = ([-0-9.e+]+), par = \(([^)]*)\)", line)
        if m:
            k = int(m.group(1))
            if k == 1 and cur:
                runs.append(cur)
                cur = []
            cur.append({"iter": k, "ssr": float(m.group(2)), "par": [float(t) for t in m.group(3).split(",")]})
            last = k
    if cur:
        runs.append(cur)
    return runs


def main():
    ut = open(os.path.join(REF, "inst/unit_tests/unit_tests_gslnls.R")).read()
    readme = open(os.path.join(REF, "README.md")).read()

    # ---- Example 1 data (given verbatim) and the RNG check --------------------------------
    x1 = parse_vector(ut, "x")
    y1 = parse_vector(ut, "y")
    rng = RRandom(1)
    z = np.array([rng.norm_rand() for _ in range(50)])
    xs = (np.arange(25)) * 3.0 / 24.0
    y1_regen = 5.0 * np.exp(-1.5 * xs) + 1.0 + 0.25 * z[:25]
    err = float(np.max(np.abs(y1_regen - np.array(y1))))
    assert np.allclose(xs, x1)
    assert err < 5e-14, "R RNG restatement does not reproduce Example-1 responses: %g" % err
    print("R RNG restatement reproduces the 25 Example-1 responses; max abs diff %.3g" % err)

    # ---- Example 2 data: set.seed(1); y = f(x) * rnorm(50, 1, 0.1) -------------------------
    x2 = np.arange(1, 51) / 50.0
    y2 = 5.0 * np.exp(-(x2 - 0.4) ** 2 / (2 * 0.15 ** 2)) * (1.0 + 0.1 * z[:50])

    traces = parse_traces(readme)
    # order in README.md: ex2a lm (26), ex2b lmaccel FD fvv (12), analytic fvv (1) (12), (2) (12)
    assert [len(t) for t in traces[:4]] == [26, 12, 12, 12], [len(t) for t in traces]

    examples = {
        "example1": {
            "source": "inst/unit_tests/unit_tests_gslnls.R:256-265; README.md:160-195,246-261",
            "formula": "y ~ A * exp(-lam * x) + b", "x": x1, "y": y1,
            "param_names": ["A", "lam", "b"], "start": [0.0, 0.0, 0.0],
            "coef_print": [4.893, 1.417, 1.010], "ssr_print": 1.316, "niter": 9,
            "finTol_print": 4.441e-16, "stderr_print": [0.1811, 0.1304, 0.1092], "sigma_print": 0.2446,
        },
        "example2": {
            "source": "README.md:470-860 (data regenerated: set.seed(1), rnorm(50, 1, 0.1))",
            "formula": "y ~ a * exp(-(x - b)^2 / (2 * c^2))", "x": x2.tolist(), "y": y2.tolist(),
            "param_names": ["a", "b", "c"], "start": [1.0, 0.0, 1.0],
            "ssr_init_print": 210.146, "ssr_final_print": 2.7583,
            "coef_print": [5.1389, 0.3979, 0.1468],
            "lm_fd": {"niter": 26, "nevalf": 124, "finTol_print": 1.33227e-15, "trace": traces[0]},
            "lmaccel_fd": {"niter": 12, "nevalf": 76, "finTol_print": 3.19744e-14, "trace": traces[1]},
            "lmaccel_fvv": {"niter": 12, "trace": traces[2]},
        },
        "example3_branin": {
            "source": "README.md:860-975",
            "start": [6.0, 14.5], "lm_coef_print": [-3.142, 12.275], "lm_niter": 20,
            "ssr_print": 0.3979, "other_methods_min": [math.pi, 2.275],
        },
        "example4_penalty": {
            "source": "README.md:977-1100", "p": 500, "alpha": 1e-5,
            "ssr_print": 0.004778845,
        },
    }
    with open(os.path.join(OUT, "readme_examples.json"), "w") as fh:
        json.dump(examples, fh, indent=1)

    probs = parse_formula_problems(os.path.join(REF, "R/nls_test.R"))
    print("parsed %d formula problems: %s" % (len(probs), ", ".join(probs)))
    with open(os.path.join(OUT, "nist_problems.json"), "w") as fh:
        json.dump({"source": "R/nls_test.R:169-990 (NIST StRD / Bates-Watts data, starts, certified values)",
                   "problems": probs}, fh, indent=1)


if __name__ == "__main__":
    main()
