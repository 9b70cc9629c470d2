"""Regenerates the golden fixtures from the reference checkout (only possible in the build
container, where /root/reference is mounted; the GPU box reads the committed outputs).

  python tests/golden/make_golden.py

Outputs (committed):
  hnsw_ingest_1000x20.f64.npy   first 1000 rows of tests/data/hnsw-random-9000-20-euclidean.gz
  hnsw_query_300x20.f64.npy     first 300 rows of tests/data/hnsw-random-5000-20-euclidean.gz
      (the row counts the reference's recall tests ingest/query: idx/trees/hnsw/mod.rs:1144-1184)
  graph_relations.json          (src, edge_table, edge_id, dst) of every RELATE in
      language-tests/tests/datasets/graph.surql, plus expected result arrays copied from
      language-tests/tests/language/graph/*.surql
"""
import gzip, json, re, os
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_gz(name, limit):
    rows = []
    with gzip.open(f"{REF}/tests/data/{name}", "rt") as fh:
        for i, line in enumerate(fh):
            if i == limit:
                break
            rows.append(json.loads(line))
    return np.asarray(rows, dtype=np.float64)


def main():
    np.save(f"{HERE}/hnsw_ingest_1000x20.f64.npy", read_gz("hnsw-random-9000-20-euclidean.gz", 1000))
    np.save(f"{HERE}/hnsw_query_300x20.f64.npy", read_gz("hnsw-random-5000-20-euclidean.gz", 300))

    rel = []
    pat = re.compile(r"RELATE\s+(\w+:\w+)->(\w+):(\w+)->(\w+:\w+)")
    for line in open(f"{REF}/language-tests/tests/datasets/graph.surql"):
        m = pat.search(line)
        if m:
            rel.append({"src": m.group(1), "edge_tb": m.group(2), "edge_id": m.group(3), "dst": m.group(4)})

    def results(fname):
        txt = open(f"{REF}/language-tests/tests/language/graph/{fname}").read()
        head = txt.split("*/")[0]
        return re.findall(r'^value = "(.*)"$', head, flags=re.M)

    def queries(fname):
        txt = open(f"{REF}/language-tests/tests/language/graph/{fname}").read()
        body = txt.split("*/", 1)[1]
        return [l.strip() for l in body.splitlines() if l.strip() and not l.strip().startswith("--")]

    cases = {}
    for f in ["traversal_multi_hop.surql", "traversal_forward.surql", "traversal_backward.surql",
              "cycles_collect.surql", "collect_min_depth.surql", "depth_fixed.surql", "depth_range.surql"]:
        cases[f] = {"results": results(f), "statements": queries(f)}
    json.dump({"relations": rel, "cases": cases}, open(f"{HERE}/graph_relations.json", "w"), indent=1)
    print(len(rel), "relations")


if __name__ == "__main__":
    main()
