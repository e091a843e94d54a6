"""Collects the GEMM shape sweeps of the reference's benchmark harness into one JSON fixture.

Source (data only, read at generation time, never at test / bench time):
    /root/reference/benchmark/config_csv/blas3/gemm/**.csv                 transA,transB,m,k,n,alpha,beta
    /root/reference/benchmark/config_csv/blas3/gemm_batched/*.csv          ...,batch_size,batch_type
    /root/reference/benchmark/config_csv/blas3/gemm_batched_strided/**.csv ...,batch_size,stride_a_mul,stride_b_mul,stride_c_mul
(column order: benchmark/README.md:118-127, common/include/common/common_utils.hpp:493-518,746-832;
note the m,k,n order).  Rows are de-duplicated per API; every row keeps the list of files it came from.

    python tests/golden/make_config_shapes.py      # writes tests/golden/config_csv_shapes.json
"""
import csv
import json
from pathlib import Path

REF = Path("/root/reference/benchmark/config_csv/blas3")
OUT = Path(__file__).resolve().parent / "config_csv_shapes.json"


def main():
    rows = {}
    for api, sub in (("gemm", "gemm"), ("gemm_batched", "gemm_batched"), ("gemm_batched_strided", "gemm_batched_strided")):
        for f in sorted((REF / sub).rglob("*.csv")):
            for r in csv.reader(open(f)):
                r = [x.strip() for x in r]
                if len(r) < 7 or r[0].lower() not in ("n", "t", "c"):
                    continue
                d = dict(api=api, ta=r[0].lower(), tb=r[1].lower(), m=int(r[2]), k=int(r[3]), n=int(r[4]),
                         alpha=float(r[5]), beta=float(r[6]))
                if api == "gemm_batched":
                    d.update(batch=int(r[7]), batch_type=r[8].lower())
                elif api == "gemm_batched_strided":
                    d.update(batch=int(r[7]), stride_a_mul=int(r[8]), stride_b_mul=int(r[9]), stride_c_mul=int(r[10]))
                key = json.dumps(d, sort_keys=True)
                rows.setdefault(key, dict(d, files=[]))["files"].append(str(f.relative_to(REF)))
    out = sorted(rows.values(), key=lambda d: (d["api"], d["m"] * d["n"] * d["k"] * d.get("batch", 1)))
    OUT.write_text(json.dumps(dict(source="codeplaysoftware/portBLAS @ 6cf5e58 benchmark/config_csv/blas3 (gemm*, data only)",
                                   rows=out), indent=0))
    print(len(out), "distinct rows ->", OUT)


if __name__ == "__main__":
    main()
