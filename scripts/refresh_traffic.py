"""Regenerate profiles/r03_traffic.json and profiles/r03_terrain_vertices_k.txt from an `ncu --set full` report of
terrain_vertices_k taken with the CURRENT terrain.cu (bench.py only quotes `roofline.traffic` when the hash matches):

    ncu --set full --import-source on --clock-control none -k regex:terrain_vertices_k -s 2 -c 1 \
        -o gpurun_out/r03_terrain_v -f python scripts/profile_terrain.py          # on the GPU box
    python scripts/refresh_traffic.py gpurun_out/r03_terrain_v.ncu-rep          # here
"""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(head)}


def metric(name):
    v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
sha = hashlib.sha256(open(os.path.join(ROOT, "myrenderer_b200", "csrc", "terrain.cu"), "rb").read()).hexdigest()
out = {"terrain_vertices_k": {
    "config": "n=4096 u16 fast32, 8-row tiles",
    "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "traffic": int(rd + wr),
    "algorithmic_bytes": 4096 * 4096 * (2 + 32),
    "terrain_cu_sha256": sha,
    "source": "profiles/r03_terrain_vertices_k.txt (ncu --set full --clock-control none, one launch of this terrain.cu)",
    "note": "below the algorithmic bytes when part of the written lines is still dirty in the 126 MB L2 at kernel end"}}
json.dump(out, open(os.path.join(ROOT, "profiles", "r03_traffic.json"), "w"), indent=1)
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "profile_summary.py"), rep, os.path.join(ROOT, "profiles", "r03_terrain_vertices_k.txt"),
                "terrain_vertices_k<u16, fast32>, n = 4096 (third launch of scripts/profile_terrain.py); terrain.cu sha256 " + sha], check=True)
print(json.dumps(out["terrain_vertices_k"]))
