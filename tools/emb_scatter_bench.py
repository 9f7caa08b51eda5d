"""Times the deterministic embedding-gradient scatter (csrc/t5misc.cu) for the token counts of 2 / 4 / 8 ranks."""
import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import t5_resnet_vqa_b200 as pkg
from t5_resnet_vqa_b200.engine import _Rec
lib = pkg.lib.load()
rec = _Rec(lib, None, lambda: torch.cuda.current_stream().cuda_stream)
D, V = 768, 32128
for world in (2, 4, 8):
    T = 2048 * world
    g = torch.Generator().manual_seed(0)
    for name, ids in (("random ids", torch.randint(2, 32100, (T,), generator=g)),
                      ("10 of 32 positions padded (id 0)", torch.where(torch.arange(T) % 32 >= 22, torch.zeros(T, dtype=torch.long), torch.randint(2, 32100, (T,), generator=g)))):
        ids = ids.cuda()
        rows = torch.randn(T, D, device="cuda")
        table = torch.zeros(V, D, device="cuda")
        first = torch.empty(2 * V, dtype=torch.int32, device="cuda")
        acc = torch.empty(T * D, dtype=torch.int64, device="cuda")
        for _ in range(3):
            rec.embedding_scatter_ordered(ids, rows, table, first, acc, T, D, V)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rec.embedding_scatter_ordered(ids, rows, table, first, acc, T, D, V)
        e1.record(); torch.cuda.synchronize()
        ref = torch.zeros(V, D, device="cuda").index_add_(0, ids, rows)
        table.zero_(); rec.embedding_scatter_ordered(ids, rows, table, first, acc, T, D, V); torch.cuda.synchronize()
        t2 = torch.zeros(V, D, device="cuda"); rec.embedding_scatter_ordered(ids, rows, t2, first, acc, T, D, V); torch.cuda.synchronize()
        assert torch.equal(t2, table)          # run-to-run identical
        print("world %d T %d %-34s %.1f us  max|diff| vs index_add %.2e" % (world, T, name, e0.elapsed_time(e1) * 100, float((table - ref).abs().max())))
