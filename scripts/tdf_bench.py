"""tsdf_ComputeTDF (separable exact distance transform) against the reference's ComputeTDF kernel rebuilt for sm_100a
(core/demo.cu:22-57) on a fragment-sized grid; bit-identity checked.  python scripts/tdf_bench.py > profiles/r02_tdf.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as graft  # noqa: E402

pkg, oracle = graft.load_package(), graft.load_oracle()
from test_tdf import shell_occupancy  # noqa: E402

peak = 6468.6
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
out = {"peak_gbs": peak, "cases": []}
for dims, voxel, trunc in [((330, 300, 260), 0.01, 0.05), ((512, 512, 512), 0.01, 0.05), ((330, 300, 260), 0.01, 0.12)]:
    occ = shell_occupancy(dims)
    n = dims[0] * dims[1] * dims[2]
    radius = int(round(trunc / voxel))
    occ_d = torch.from_numpy(occ).cuda()
    tdf_d = torch.zeros_like(occ_d)
    stream = torch.cuda.current_stream()
    pkg.ComputeTDF(occ_d, tdf_d, *dims, 0.0, 0.0, 0.0, voxel, trunc, stream=stream.cuda_stream)
    times = []
    for _ in range(5):
        tdf_d.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pkg.ComputeTDF(occ_d, tdf_d, *dims, 0.0, 0.0, 0.0, voxel, trunc, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    bpv = 4 + 4 + 8 + (4 if 3 * radius * radius < 255 else 8)   # occ twice, tdf read + write, two intermediates written + read
    case = {"dims": dims, "radius": radius, "ms": round(ms, 4), "algorithmic_bytes_per_voxel": bpv,
            "achieved_gbs": round(n * bpv / ms / 1e6, 1), "frac_of_peak": round(n * bpv / ms / 1e6 / peak, 3)}
    if oracle.REF_TDF_CUDA_LIB.exists() and n < 2 ** 27:
        ref, ref_ms = oracle.ref_cuda_compute_tdf(occ, dims, voxel, trunc)
        case["reference_kernel_ms"] = round(ref_ms, 3)
        case["speedup_vs_reference_kernel"] = round(ref_ms / ms, 1)
        case["bit_identical_to_reference_kernel"] = bool(np.array_equal(tdf_d.cpu().numpy().view(np.uint32), ref.view(np.uint32)))
    out["cases"].append(case)
print(json.dumps(out))
