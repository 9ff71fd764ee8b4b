import sys, os
sys.path.insert(0, "/root/repo")
import bench, mmc_b200
from mmc_b200 import api
for name, method in (("sphshells", "grid"), ("cube60", "elem")):
    cfg, desc = bench.workload(name, method)
    cfg["nphoton"] = 10000000
    for env in ("", "noscout", "nohot"):
        os.environ.pop("MMCB_NO_SCOUT", None)
        c = dict(cfg)
        if env == "noscout":
            os.environ["MMCB_NO_SCOUT"] = "1"
        if env == "nohot":
            c["hotcache"] = -1
        s = mmc_b200.Session(c)
        ms = []
        for i in range(4):
            s.launch(10000000, photon_offset=0, seed=cfg["seed"], seed_offset=i)
            ms.append(round(s.sync(), 2))
        s.close()
        print(name, method, env or "default", ms, flush=True)
