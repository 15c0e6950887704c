"""Dev helper: the few numbers of a bench.py JSON line that the docs quote."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g pts/s, %.4f ms/step (warm %.4f); e2e %.4g pts/s, %.4f ms (pageable %.4f)" % (d["value"], d["ms_per_step"], d["ms_per_update_warm_l2"],
      d["e2e"]["value"], d["e2e"]["ms_per_update"], d["e2e"].get("pageable_input_ms_per_update", float("nan"))))
print("clocks", d["clocks"], "launches", d["gpu_launches"])
print("roofline k_search frac %.4f (%.1f us), batched %.4f (%.3f ms); shares %s" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"] * 1e3,
      d["roofline"]["batched"]["frac"], d["roofline"]["batched"]["kernel_ms"], d["kernels"]["share_of_step"]))
for k in ("reloc", "ndt", "fullmap", "sequence", "scan2map"):
    if k not in d:
        continue
    v = d[k]
    if k == "reloc":
        print("reloc %.4g hyp/s, %.3f ms/batch" % (v["value"], v["ms_per_batch"]), v.get("parity"))
    elif k == "ndt":
        print("ndt set_target %.3f ms, derivatives %.4f ms, align %.4f ms (e2e %.4f), launches %d" % (v["set_target_ms"]["device"], v["derivatives_ms"]["device"],
              v["align_ms"]["device"], v["align_ms"]["e2e_wall"], v["align_ms"]["launches"]), v.get("parity"))
    elif k == "fullmap":
        print("fullmap %.4g keyframes/s (%.1f ms), e2e %.4g" % (v["value"], v["seconds"] * 1e3, v["e2e"]["value"]), v.get("parity"))
    elif k == "sequence":
        print("sequence update mean %.3f ms median %.3f, per scan e2e %.3f ms (MapIncremental %.3f)" % (v["ms_update_device"]["mean"], v["ms_update_device_median"],
              v["ms_per_scan_e2e"]["mean"], v["ms_per_scan_e2e"]["map_incremental_mean"]), v["parity_vs_oracle"]["max_state_diff"])
    elif k == "scan2map":
        print("scan2map optimize %.3f ms (e2e %.3f), iters %d" % (v["optimize_ms"]["device"], v["optimize_ms"]["e2e_wall"], v["iters"]), v.get("parity"))
print("cpu_baseline", d.get("cpu_baseline")); print("parity", d.get("parity"))
