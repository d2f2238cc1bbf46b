#!/usr/bin/env python
"""Timeline probes (measurement aids): clock stamps of the screen kernel's MMA / epilogue hand-off for
CTA 0's first item, and phase timestamps of the fused ladder kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200._lib import lib, ptr
from tscode_b200.rmsd_pruning import RmsdPruner
from tscode_b200.synth import gen_ensemble
S = gen_ensemble(3, 50000, 80, 5000)
pr = RmsdPruner(S, np.full(80, 6), 0.5, variant=(sys.argv[1] if len(sys.argv) > 1 else "f16"),
                grid_ctas=(int(sys.argv[2]) if len(sys.argv) > 2 else 0))
trace = torch.zeros(8 * 96, dtype=torch.int64, device="cuda")
pr.run(); torch.cuda.synchronize()
lib().tsc_set_trace_buffer(ptr(trace))
pr.pack(); pr.screen(); torch.cuda.synchronize()
lib().tsc_set_trace_buffer(None)
t = trace.cpu().numpy().reshape(96, 8)
t0 = t[0, 0]
print("tile: mma_start b_full t_empty issued | epi_start t_full done   (cycles since tile 0 mma_start)")
for k in range(40, 72):
    print(k, *(int(x - t0) for x in t[k, :7]))
d = np.diff(t[32:96, 3]); print("issue-to-issue cycles/tile: mean", d.mean(), "min", d.min(), "max", d.max())
print("mma wait b_full", (t[32:96, 1] - t[32:96, 0]).mean(), "wait t_empty", (t[32:96, 2] - t[32:96, 1]).mean(),
      "issue", (t[32:96, 3] - t[32:96, 2]).mean())
print("epi: wait t_full", (t[32:96, 5] - t[32:96, 4]).mean(), "work", (t[32:96, 6] - t[32:96, 5]).mean())
print("issued -> t_full seen", (t[32:96, 5] - t[32:96, 3]).mean())
print("SM clock from clock64 / globaltimer over tiles 8..95: %.1f MHz" % (1e3 * (t[95, 3] - t[8, 3]) / max(t[95, 7] - t[8, 7], 1)))
pr.verify(); m = pr.eliminate(); torch.cuda.synchronize()
info = pr.fused_out[pr._info_off:pr._info_off + 256].view(torch.int32).cpu().numpy()
print("fused ladder status", info[:4], "rounds", info[8:8 + info[1]])
print("fused ladder stamps (ns):", info[32:62].tolist())
