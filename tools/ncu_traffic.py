"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration of the
flat-step kernels from `ncu --set full` reports into profiles/r01_ncu_traffic.json; bench.py reads it to
fill `roofline.traffic` for the dominant kernel."""
import csv, json, subprocess, sys
LABELS = [("EpiHeadNorm", "K2_head_proj_norm_fwd_cluster"), ("EpiAtomicAddF32, 0, 0>", "K2_head_splitk_gemm"),
          ("bias_norm_rows", "K2_head_bias_norm"), ("text_encoder_flat_wide", "K1_text_encoder_fwd"),
          ("eval_nway_stream", "K7_eval_nway_stream"),
          ("EpiSimStats", "K3K4_sim_infonce_fwd"), ("EpiGradG", "K5a_sim_infonce_bwd_g"),
          ("EpiNormBwdT<0>", "K5b_dimg_norm_bwd"), ("EpiNormBwdT<1>", "K5b_dtxt_norm_bwd"),
          ("EpiStoreF32", "K5c_head_weight_grad"), ("text_encoder_fwd", "K1_text_encoder_fwd"),
          ("embedding_scatter_add", "K5e_embedding_scatter_add"), ("cast_f32_bf16", "cast_w_f32_to_bf16"),
          ("infonce_finalize", "infonce_finalize")]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}
out = {}
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    def val(r, name):
        i = h.index(name)
        return float(r[i]) * UNIT.get(u[i], 1)
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        for key, label in LABELS:
            if key in name and label not in out:
                out[label] = dict(kernel=name[:120], dram_bytes_read=val(r, "dram__bytes_read.sum"),
                                  dram_bytes_write=val(r, "dram__bytes_write.sum"),
                                  duration_s=val(r, "gpu__time_duration.sum"), report=rep.split("/")[-1])
                out[label]["traffic"] = out[label]["dram_bytes_read"] + out[label]["dram_bytes_write"]
if "K2_head_splitk_gemm" in out and "K2_head_bias_norm" in out:      # the head as bench.py times it: both kernels
    a, b = out["K2_head_splitk_gemm"], out["K2_head_bias_norm"]
    out["K2_head_proj_norm_fwd"] = dict(kernel=a["kernel"] + " + " + b["kernel"], traffic=a["traffic"] + b["traffic"],
                                        dram_bytes_read=a["dram_bytes_read"] + b["dram_bytes_read"],
                                        dram_bytes_write=a["dram_bytes_write"] + b["dram_bytes_write"],
                                        duration_s=a["duration_s"] + b["duration_s"], report=a["report"])
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: (v["traffic"], v["duration_s"]) for k, v in out.items()}, indent=0))
