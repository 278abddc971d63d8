"""Builds tests/golden/tpims_links.npz from the reference's own link tables.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_tpims_fixture.py
Source: /root/reference/dataset/tpims_link_0322.tar.xz -> link_{IA,KS,KY,OH,WI}_data.csv with
columns SRC_IDX,SRC,DST_IDX,DST,DIST (load_dataset.py:303-305); node ids are the SRC_IDX/DST_IDX
columns (the reference's ``mapping`` at load_dataset.py:336 is the identity on them);
num_nodes = rows of dataset/data/tpims_location.csv minus the IL/MI/MN/IN sites (load_dataset.py:333).
"""
import csv, io, os, tarfile
import numpy as np

REF = "/root/reference"
src, dst, dist, region = [], [], [], []
with tarfile.open(os.path.join(REF, "dataset", "tpims_link_0322.tar.xz")) as tf:
    members = {os.path.basename(m.name): m for m in tf.getmembers() if m.isfile()}
    for r, st in enumerate(["IA", "KS", "KY", "OH", "WI"]):
        f = io.TextIOWrapper(tf.extractfile(members[f"link_{st}_data.csv"]))
        for row in csv.reader(f):
            src.append(int(row[0])); dst.append(int(row[2])); dist.append(float(row[4])); region.append(r)
with open(os.path.join(REF, "dataset", "data", "tpims_location.csv")) as f:
    rows = list(csv.DictReader(f))
n = sum(1 for r in rows if not r["SITE_ID"].startswith(("IL", "MI", "MN", "IN")))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tpims_links.npz")
np.savez(out, src=np.array(src, np.int64), dst=np.array(dst, np.int64), dist=np.array(dist, np.float32),
         region=np.array(region, np.int32), num_nodes=np.int64(n))
print(out, len(src), "edges", n, "nodes", "max id", max(max(src), max(dst)))
