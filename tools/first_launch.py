import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, mmc_b200
from mmc_b200 import api
for name, method in (("sphshells", "grid"), ("sphshells", "elem"), ("cube60", "elem"), ("headatlas", "elem")):
    cfg, desc = bench.workload(name, method)
    cfg["nphoton"] = 10000000
    for env in ("", "", "nohot"):       # the first row of a workload also pays the lazy load of the scout's kernel
        os.environ.pop("MMCB_NO_SCOUT", None)
        c = dict(cfg)
        if env == "noscout":
            os.environ["MMCB_NO_SCOUT"] = "1"
        if env == "nohot":
            c["hotcache"] = -1
        s = mmc_b200.Session(c)
        ms = []
        for i in range(3):
            s.launch(10000000, photon_offset=0, seed=cfg["seed"], seed_offset=i)
            ms.append(round(s.sync(), 2))
        s.fetch()
        s.close()
        print(name, method, env or "default", ms, flush=True)
