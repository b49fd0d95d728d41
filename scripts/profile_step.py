"""One warm-up + one device-resident step of the C2 hot path (sort old, search new) for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deltaq_b200 import CudaSuffixSort, workloads as w  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
old, new = w.c2_exe_pair()
n, m = old.size, new.size
s = CudaSuffixSort()
ctx = s.context
d_old = torch.from_numpy(old).cuda()
d_new = torch.from_numpy(new).cuda()
d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
d_pos = torch.empty(m, dtype=torch.int32, device="cuda")
d_len = torch.empty(m, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(steps):
    ctx.suffix_sort_device(d_old.data_ptr(), n, d_sa.data_ptr())
    ctx.bsdiff_search_device(d_old.data_ptr(), n, None, d_new.data_ptr(), m, 0, m, d_pos.data_ptr(), d_len.data_ptr())
print(ctx.stats())
