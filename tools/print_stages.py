import json,sys
j=json.loads(sys.stdin.read()); print(sys.argv[1], round(j["value"],1), {k:v["ms"] for k,v in j["stages"].items() if k in ("presort","sort","duplicate","scan")})
