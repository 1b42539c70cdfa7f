import json, sys
d = json.loads(sys.stdin.read())
e = d["e2e"]
print(sys.argv[1], "device", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms | e2e", round(e["value"], 1),
      "| blocking ms", round((e.get("blocking") or {}).get("ms_per_frame", 0), 3), "|", {k: round(v, 3) for k, v in d["kernel_ms_per_frame"].items() if isinstance(v, (int, float))}, "| latency ms", round(d.get("frame_latency_ms") or 0, 3), "| enqueue ms", (d.get("run") or d["config"]).get("host_enqueue_ms_per_frame"))
