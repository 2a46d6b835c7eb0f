"""Pin the oracle against the reference itself and (re)generate tests/golden/.

Runs ONLY in the build container (needs /root/reference, which does not exist on the GPU box):

    python oracle/pin_against_reference.py [--ref /root/reference] [--write]

What it does
  1. imports the reference's own Python module FBTT/tt_embeddings_ops.py (with an empty stub for
     the compiled `tt_embeddings` extension it imports at module top) and uses
       * tt_matrix_to_full(p, q, ranks, cores, [1, 0, 2, 3])   (:80-127, the pure-PyTorch TT
         contraction) + F.embedding_bag(mode="sum", include_last_offset=True) as the forward
         truth  -- the assertion the reference's own test states at sage_profiler.py:303-305
       * autograd through the same expression as the dense-gradient truth (:362-367)
       * suggested_tt_shapes (:369-429)
  2. compiles a 20-line host program that includes the reference's FBTT/hashtbl_cuda_utils.cuh
     from where it lies and prints murmor_hash_3_32 for a key list (known-answer vectors)
  3. checks oracle/tt_oracle.c against all of it, and with --write stores the vectors as small
     fixtures under tests/golden/ so the parity tests can run where the reference is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

CASES = [
    # name, num_embeddings, p, q, ranks, n_bags, bagged
    ("cora_r16", 2708, [14, 14, 14], [4, 4, 8], [16, 16], 300, False),
    ("cora_r16_bags", 2708, [14, 14, 14], [4, 4, 8], [16, 16], 64, True),
    ("products_small", 5 * 6 * 7, [5, 6, 7], [4, 5, 5], [16, 16], 150, False),
    ("products_small_bags", 5 * 6 * 7, [5, 6, 7], [4, 5, 5], [16, 16], 48, True),
    ("papers_small_r32", 4 * 5 * 6, [4, 5, 6], [4, 4, 8], [32, 32], 100, False),
    ("two_cores", 9 * 11, [9, 11], [4, 8], [12], 80, True),
    ("four_cores", 3 * 4 * 5 * 6, [3, 4, 5, 6], [2, 2, 3, 4], [4, 6, 5], 90, True),
]

HASH_KEYS = [0, 1, 2, 139, 140, 17499, 12345, 196614, 2449028, 111059955, (1 << 32) + 7,
             (1 << 40) + 12345, 2449029 * 3 + 1, 7, 1023, 65536, 999983]
HASH_SIZES = [2449029, 169343, 2708, 111059956, 97, 1000]


def import_reference(ref):
    sys.modules.setdefault("tt_embeddings", types.ModuleType("tt_embeddings"))
    sys.path.insert(0, ref)
    from FBTT import tt_embeddings_ops as ops  # noqa
    return ops


def reference_hashes(ref):
    """Compile the reference header on the host and evaluate its hash."""
    import torch.utils.cpp_extension as ce
    src = r"""
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "%s/FBTT/hashtbl_cuda_utils.cuh"
int main(int argc, char** argv) {
  for (int i = 1; i + 1 < argc; i += 2) {
    long long key = atoll(argv[i]); int size = atoi(argv[i + 1]);
    printf("%%lld %%d %%u\n", key, size, murmor_hash_3_32((int64_t)key, (int32_t)size));
  }
  return 0;
}
""" % ref
    tmp = tempfile.mkdtemp(prefix="refhash")
    cu = os.path.join(tmp, "h.cu")
    exe = os.path.join(tmp, "h")
    with open(cu, "w") as f:
        f.write(src)
    inc = []
    for p in ce.include_paths():
        inc += ["-I", p]
    cmd = ["nvcc", "-std=c++17", "--expt-relaxed-constexpr", "-O1", "-o", exe, cu] + inc + [
        "-L", os.path.join(os.path.dirname(torch.__file__), "lib"), "-lc10", "-ltorch_cpu",
        "-Xlinker", "-rpath," + os.path.join(os.path.dirname(torch.__file__), "lib")]
    subprocess.run(cmd, check=True, capture_output=True)
    args = []
    for s in HASH_SIZES:
        for k in HASH_KEYS:
            args += [str(k), str(s)]
    out = subprocess.run([exe] + args, check=True, capture_output=True, text=True).stdout
    res = []
    for line in out.strip().splitlines():
        k, s, h = line.split()
        res.append([int(k), int(s), int(h)])
    return res


def make_case(ops, name, n_emb, p, q, ranks, n_bags, bagged, seed):
    g = torch.Generator().manual_seed(seed)
    T = len(p)
    rr = [1] + list(ranks) + [1]
    D = int(np.prod(q))
    cores = [torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) * (0.5 / np.sqrt(rr[t]))
             for t in range(T)]
    rng = np.random.default_rng(seed)
    if bagged:  # pooling factors like sage_profiler.py:71-100 (N(10, 20^2) clipped at 0)
        lengths = np.clip(np.round(rng.normal(3.0, 4.0, size=n_bags)), 0, None).astype(np.int64)
        lengths[0] = 0          # an empty bag at the front
        lengths[-1] = 0         # and at the back
    else:
        lengths = np.ones(n_bags, dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    nnz = int(offsets[-1])
    indices = rng.integers(0, n_emb, size=nnz).astype(np.int64)
    if nnz > 4:               # make sure duplicates and the extreme rows are exercised
        indices[0] = 0
        indices[1] = n_emb - 1
        indices[2] = indices[3]
    d_out = (torch.rand(n_bags, D, generator=g) * 0.1)

    # --- reference forward in fp32 (what the reference module's full_weight() gives)
    W32 = ops.tt_matrix_to_full(p, q, rr, cores, [1, 0, 2, 3])
    out32 = torch.nn.functional.embedding_bag(torch.from_numpy(indices), W32,
                                              torch.from_numpy(offsets), mode="sum",
                                              include_last_offset=True)
    # --- the same in float64 + autograd for the dense gradients
    cores64 = [c.double().requires_grad_(True) for c in cores]
    # tt_matrix_to_full ends with .float(); keep double by undoing the cast's rounding effect:
    # run it on double inputs and use the result before the cast via a local re-implementation
    # check (the cast only rounds the values).
    W64f = ops.tt_matrix_to_full(p, q, rr, cores64, [1, 0, 2, 3])  # float32 view, grads in double
    out64 = torch.nn.functional.embedding_bag(torch.from_numpy(indices), W64f.double(),
                                              torch.from_numpy(offsets), mode="sum",
                                              include_last_offset=True)
    (out64 * d_out.double()).sum().backward()
    d_cores = [c.grad.float().numpy() for c in cores64]
    rowidx = np.repeat(np.arange(n_bags, dtype=np.int64), lengths)
    return {
        "p": np.array(p), "q": np.array(q), "ranks": np.array(rr), "num_embeddings": n_emb,
        "indices": indices, "offsets": offsets, "rowidx": rowidx,
        "d_output": d_out.numpy(), "out": out32.numpy(), "out64": out64.detach().float().numpy(),
        **{"core%d" % t: cores[t].numpy() for t in range(T)},
        **{"d_core%d" % t: d_cores[t] for t in range(T)},
    }


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    ops = import_reference(args.ref)
    ok = True

    # 1. forward / backward truth
    bundle = {}
    for i, (name, n_emb, p, q, ranks, n_bags, bagged) in enumerate(CASES):
        c = make_case(ops, name, n_emb, p, q, ranks, n_bags, bagged, seed=100 + i)
        T = len(p)
        cores = [c["core%d" % t] for t in range(T)]
        out = orc.tt_forward(p, q, list(c["ranks"]), cores, c["indices"], c["rowidx"], n_bags)[0]
        e_f = rel_err(out, c["out64"])
        dc = orc.tt_backward_dense(p, q, list(c["ranks"]), cores, c["indices"], c["rowidx"],
                                   c["d_output"][None])
        e_b = max(rel_err(dc[t], c["d_core%d" % t]) for t in range(T))
        print("%-22s fwd rel err %.2e (fp32 ref path %.2e)  bwd rel err %.2e" %
              (name, e_f, rel_err(c["out"], c["out64"]), e_b))
        ok &= e_f < 2e-6 and e_b < 2e-6
        for k, v in c.items():
            bundle[name + "/" + k] = v

    # 2. row layout: reference full matrix rows vs oracle single rows
    p, q, rr = [14, 14, 14], [4, 4, 8], [1, 16, 16, 1]
    g = torch.Generator().manual_seed(7)
    cores = [torch.randn(1, p[t], rr[t] * q[t] * rr[t + 1], generator=g) for t in range(3)]
    W = ops.tt_matrix_to_full(p, q, rr, cores, [1, 0, 2, 3]).numpy()
    rows = np.array([0, 1, 13, 14, 195, 196, 2707, 2743], dtype=np.int64)
    o = orc.tt_forward(p, q, rr, [c.numpy() for c in cores], rows, np.arange(rows.size), rows.size)[0]
    e = rel_err(o, W[rows])
    print("row layout check: rel err %.2e" % e)
    ok &= e < 2e-6
    bundle["rowlayout/rows"] = rows
    bundle["rowlayout/values"] = W[rows]
    for t in range(3):
        bundle["rowlayout/core%d" % t] = cores[t].numpy()

    # 3. suggested_tt_shapes known answers
    shapes = {}
    try:
        shapes = {"2708_3": ops.suggested_tt_shapes(2708, 3),
                  "128_3_noround": ops.suggested_tt_shapes(128, 3, allow_round_up=False),
                  "100_3_noround": ops.suggested_tt_shapes(100, 3, allow_round_up=False),
                  "169343_3": ops.suggested_tt_shapes(169343, 3)}
        shapes = {k: [int(x) for x in v] for k, v in shapes.items()}
        print("suggested_tt_shapes:", shapes)
    except Exception as ex:  # sympy/scipy absent
        print("suggested_tt_shapes skipped:", ex)

    # 4. hash known answers from the reference header
    kat = reference_hashes(args.ref)
    bad = [(k, s, h, orc.hash32(k, s)) for k, s, h in kat if orc.hash32(k, s) != h]
    print("hash KAT: %d vectors, %d mismatches" % (len(kat), len(bad)))
    ok &= not bad

    # 5. Efficient_TT float index math == integer math on the shapes it is defined for
    for name, N, pp in [("cora", 2708, [14, 14, 14]), ("arxiv", 169343, [55, 55, 56]),
                        ("products", 2449029, [125, 140, 140])]:
        idx = np.arange(N, dtype=np.int64)
        same = np.array_equal(orc.eff_split(idx, pp, True), orc.eff_split(idx, pp, False))
        print("eff float==int on %s: %s" % (name, same))
        ok &= same

    if args.write:
        gdir = os.path.join(ROOT, "tests", "golden")
        os.makedirs(gdir, exist_ok=True)
        np.savez_compressed(os.path.join(gdir, "tt_cases.npz"), **bundle)
        with open(os.path.join(gdir, "hash_kat.json"), "w") as f:
            json.dump({"source": "murmor_hash_3_32 compiled from FBTT/hashtbl_cuda_utils.cuh:48-76",
                       "vectors": kat, "suggested_tt_shapes": shapes}, f)
        print("wrote", gdir)
    print("PIN", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
