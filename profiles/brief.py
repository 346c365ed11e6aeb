import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["stage_ms_per_step"])
