python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from radlite_b200 import synth
from radlite_b200.api import Renderer
for n in (3, 1):
    m = synth.config(n)
    g = Renderer(0)
    g.load_model(m)
    for it in range(8):
        if it >= 4:
            g.invalidate_geometry()
        ms = g.render_device(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
        print("cfg", n, "it", it, "rebuild" if it >= 4 else "cached", [round(x, 2) for x in ms], flush=True)
    g.close()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --config 3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
python scripts/launch_shares.py gpurun_out/launches_c3.csv | head -12
grep tile_kernel gpurun_out/launches_c3.csv | awk -F, '{print $NF}'
