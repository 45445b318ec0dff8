#!/usr/bin/env python
"""profiles/r02_dram_traffic.json from an `ncu --set full` capture of one build (tools/profile_build.py):
dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the kernel labels bench.py uses in roofline.kernels.
  python tools/dram_traffic.py gpurun_out/r02_final.ncu-rep --scale 1.0 --variant clicks > profiles/r02_dram_traffic.json"""
import argparse, csv, io, json, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--variant", default="clicks")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
label = [("pairgen_kernel<0>", "pairgen_kernel<count>"), ("pairgen_kernel<1>", "pairgen_kernel<scatter>"),
         ("partition_kernel<0>", "partition_kernel<count> (+ bin offsets scan)"), ("partition_kernel<1>", "partition_kernel<move>"),
         ("otable_warp_kernel", "reduce_classify + otable_warp_kernel (bins <= 384 records)"),
         ("otable_block_kernel<1, 512", "otable_block_kernel<512 threads> (bins <= 6144)"),
         ("otable_block_kernel<1, 256", "otable_block_kernel<256 threads> (bins <= 3072)"),
         ("otable_block_kernel<1, 128", "otable_block_kernel<128 threads> (bins <= 1536)"),
         ("otable_block_kernel<0, 512", "otable_block_kernel<512 threads> (bins <= 6144)"),
         ("otable_block_kernel<0, 256", "otable_block_kernel<256 threads> (bins <= 3072)"),
         ("otable_block_kernel<0, 128", "otable_block_kernel<128 threads> (bins <= 1536)"),
         ("tail_copy_all_kernel", "tail_copy_all_kernel")]
out = {}
for r in rows[2:]:
    name = r[ki]
    for pat, lab in label:
        if pat in name:
            b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            out[lab] = out.get(lab, 0) + int(b)
print(json.dumps({"source": f"{a.rep} (ncu --set full --clock-control none, python tools/profile_build.py --scale {a.scale:g}, one launch each)",
                  "scale": a.scale, "variant": a.variant, "unit": "bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum",
                  "kernels": out}, indent=1))
