"""Record the DRAM bytes of one kernel launch (from a profiles/*.json summary written by tools/ncu_summary.py) in profiles/traffic.json
together with a hash of the kernel sources, so that bench.py only quotes it while the sources are unchanged.
Usage: python tools/record_traffic.py profiles/<summary>.json "<bench key: kernel_NxM_bB_S>" """
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    summary, key = sys.argv[1], sys.argv[2]
    d = json.load(open(summary))
    val = lambda k: float(str(d[k]["value"]).replace(",", ""))
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = sum(val(k) * unit[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    t = {k: v for k, v in t.items() if isinstance(v, dict)}  # drop the pre-hash format
    t[key] = {"bytes": total, "csrc": bench.csrc_hash(key), "capture": os.path.basename(summary)}
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)
    print(key, total, "bytes per launch; csrc", bench.csrc_hash(key))


if __name__ == "__main__":
    main()
