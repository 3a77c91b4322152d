#!/bin/bash
PARSEC_MCA_device_b200_trace=/tmp/chain PARSEC_MCA_device_b200_enabled=1 PARSEC_MCA_device_b200_memory_number_of_blocks=1024 timeout 60 oracle/_ref/bin/ex02_b200 -m gpu -N 999 -c 2 -r 2 2>&1 | tail -1 | cut -c150-330
python - <<'PY'
import json, glob
ev = json.load(open(glob.glob("/tmp/chain.*.json")[0]))["traceEvents"]
ev = ev[-1000:]                      # the last repeat
ev.sort(key=lambda e: e["ts"])
dur = sorted(e["dur"] for e in ev)
gap = sorted(ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]) for i in range(len(ev) - 1))
period = sorted(ev[i + 1]["ts"] - ev[i]["ts"] for i in range(len(ev) - 1))
med = lambda a: a[len(a) // 2]
print("tasks", len(ev), "in-kernel task time us: median %.2f p90 %.2f" % (med(dur), dur[int(len(dur) * .9)]))
print("gap end(k) -> start(k+1) us: median %.2f p10 %.2f p90 %.2f" % (med(gap), gap[int(len(gap) * .1)], gap[int(len(gap) * .9)]))
print("period us: median %.2f" % med(period))
PY
