"""FP32 (no FMA) and L2 read ceilings of this box -> gpurun_out/measured_fp32_l2_peaks.json (copied to profiles/)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("cpp-11-ray-trace-march-framework_b200.capi")
runs = [capi.measure_peaks(0) for _ in range(3)]
out = {"fp32_nonfma_tflops": max(r[0] for r in runs), "l2_read_gbps": max(r[1] for r in runs), "runs": runs,
       "nominal_fp32_nonfma_tflops": 148 * 128 * 1.965e9 / 1e12,
       "method": "csrc/measure.cu: 8 independent FMUL+FADD chains per thread, 2 x 1024 threads per SM; every CTA streams a 48 MB "
                 "buffer with 16-byte ld.global.cg loads (L2-resident), best of 3 timed launches"}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/measured_fp32_l2_peaks.json", "w"), indent=1)
print(json.dumps(out))
