"""lm_head search kernel (512 MB tied embedding, 256 rows): TMA-box operands vs packed tile images + bulk copies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mic_b200
from mic_b200 import ops, generation as gen

dev, bf = "cuda", torch.bfloat16
R, d, V = 256, 1024, 250054
h = (torch.randn(R, d, device=dev) * 0.5).to(bf)
E = (torch.randn(V, d, device=dev) * 0.05).to(bf)
bias = torch.zeros(V, device=dev)
n = ops.lm_head_search_num_partials(R)
ws = {"nparts": n, "pmax": torch.empty(n, R, device=dev), "psum": torch.empty(n, R, device=dev),
      "cand_val": torch.empty(n, R, 8, device=dev), "cand_idx": torch.empty(n, R, 8, device=dev, dtype=torch.int32),
      "row_lp": torch.empty(R, 8, device=dev), "row_tok": torch.empty(R, 8, device=dev, dtype=torch.int32),
      "row_ml": torch.empty(R, 2, device=dev)}
def aligned(nbytes):
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    off = (-raw.data_ptr()) % 1024
    return raw[off:off + nbytes]
ht = aligned(ops.pack_kmajor_tiles_bytes(R, d, 128)); et = aligned(ops.pack_kmajor_tiles_bytes(V, d, 256))
ops.pack_kmajor_tiles(h, 128, ht); ops.pack_kmajor_tiles(E, 256, et)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
ops.lm_head_search(h, E, bias, -1, ws); ops.search_merge(ws, R); torch.cuda.synchronize()
a_lp, a_tok = ws["row_lp"].clone(), ws["row_tok"].clone()
ops.lm_head_search_packed(ht, et, bias, -1, R, V, d, ws); ops.search_merge(ws, R); torch.cuda.synchronize()
print("same tokens:", torch.equal(a_tok, ws["row_tok"]), " max lp diff:", (a_lp - ws["row_lp"]).abs().max().item())
print(f"search (TMA boxes) : {timeit(lambda: ops.lm_head_search(h, E, bias, -1, ws)):8.1f} us")
print(f"search (packed)    : {timeit(lambda: ops.lm_head_search_packed(ht, et, bias, -1, R, V, d, ws)):8.1f} us")
print(f"merge              : {timeit(lambda: ops.search_merge(ws, R)):8.1f} us")
print(f"pack embedding     : {timeit(lambda: ops.pack_kmajor_tiles(E, 256, et), 5):8.1f} us")
# ---- what bounds it?  Same operand stream with (a) half the rows (one m-tile: every weight tile is used once), (b) a plain
# bf16-store epilogue instead of the search epilogue, (c) R = 128 with the search epilogue
Vp = 250048
for rows in (256, 128):
    hh = h[:rows].contiguous()
    out = torch.empty(rows, Vp, dtype=bf, device=dev)
    t = timeit(lambda: ops.gemm(hh, E[:Vp], out=out, block_n=256))
    print(f"plain GEMM [{rows} x {Vp} x {d}] bf16 store : {t:8.1f} us  ({2.0 * rows * Vp * d / t / 1e6:6.0f} TFLOP/s, {Vp * d * 2 / t / 1e3:6.0f} GB/s of weights)")
n128 = ops.lm_head_search_num_partials(128)
ws128 = {"nparts": n128, "pmax": torch.empty(n128, 128, device=dev), "psum": torch.empty(n128, 128, device=dev),
         "cand_val": torch.empty(n128, 128, 8, device=dev), "cand_idx": torch.empty(n128, 128, 8, device=dev, dtype=torch.int32),
         "row_lp": torch.empty(128, 8, device=dev), "row_tok": torch.empty(128, 8, device=dev, dtype=torch.int32),
         "row_ml": torch.empty(128, 2, device=dev)}
h128 = h[:128].contiguous()
t = timeit(lambda: ops.lm_head_search(h128, E, bias, -1, ws128))
print(f"search (TMA boxes), 128 rows        : {t:8.1f} us  ({V * d * 2 / t / 1e3:6.0f} GB/s of weights)")
