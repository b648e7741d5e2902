#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line (headline, e2e, every `extra` entry)."""
import json
import sys


def main():
    for f in sys.argv[1:]:
        try:
            d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        except Exception as e:  # noqa: BLE001
            print(f, "NO JSON", e)
            continue
        r = d.get("roofline", {})
        print("%s: n=%s mode=%s K=%s value=%.4g us/step=%.3f frac=%.3f launches=%s per_rank_ms=%s" % (
            f.split("/")[-1], d.get("n_gpus"), d.get("launch_config", d.get("config", {})).get("mode"), d.get("steps"), d["value"],
            d["ms_per_step"] * 1e3, r.get("frac", float("nan")), d.get("gpu_launches"),
            ["%.3f" % x for x in d.get("per_rank_ms", [])]))
        c = d.get("clocks") or {}
        print("   clocks: sm %s / %s MHz reasons=%s" % (c.get("sm_mhz"), c.get("sm_max_mhz"), c.get("reasons")))
        e = d.get("e2e")
        if e:
            print("   e2e %.4g (%s B D2H/step, %.1f ms/step)%s" % (e["value"], e.get("d2h_bytes_per_step"), e.get("ms_per_step", 0),
                  (("  unpipelined fp32 %.4g" % e["unpipelined_fp32"]["value"]) if "unpipelined_fp32" in e else "") +
                  (("  packed u2 %.4g" % e["packed_u2_tiles"]["value"]) if "packed_u2_tiles" in e else "") +
                  (("  int8 tiles %.4g" % e["int8_tiles"]["value"]) if "int8_tiles" in e else "")))
        if "cpu_baseline" in d:
            print("   cpu_baseline %.4g (%s cores)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
        x = d.get("extra", {})
        for k in ("per_step_launches", "per_step_stream_ordered", "fused_rollout_T33_philox", "fused_int8_tiles", "fused_u2_tiles"):
            if k in x:
                print("   %-26s %.4g  %.3f us/step  frac %.3f" % (k, x[k]["value"], x[k].get("us_per_step", x[k].get("ms_per_launch", 0) * 1e3 / 33), x[k]["frac"]))
        for wl, w in x.get("workloads", {}).items():
            parts = []
            for k in ("fused", "per_step_chained", "per_step_stream_ordered"):
                if k in w:
                    parts.append("%s %.4g (%.3f us, %.3f)" % (k, w[k]["value"], w[k]["us_per_step"], w[k]["frac"]))
            print("   %-4s %s" % (wl, "; ".join(parts)))
        for k, v in x.items():
            if k.startswith("gather_"):
                print("   %-28s %s" % (k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "per_rank_ms"}))
        if "dropin_b1" in x:
            print("   dropin_b1", {a: (round(b, 2) if isinstance(b, float) else b) for a, b in x["dropin_b1"].items() if a != "note"})
        for line in x.get("sweep", []):
            print("   sweep %-4s B=%-8d %s" % (line["workload"], line["envs_per_gpu"], "; ".join(
                "%s %.4g (%.3f)" % (k, line[k]["value"], line[k]["frac"]) for k in ("fused", "per_step_chained") if k in line)))


if __name__ == "__main__":
    main()
