"""Developer check run on a GPU box: product gen_data vs the oracle port on fresh synthetic sequences.
(Lives under tests/: only test code may import the oracle.)"""
import filecmp, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voxelizer_b200 as vx
from oracle.oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def compare(a, b):
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    if fa != fb:
        return [f"file lists differ: {len(fa)} vs {len(fb)}"]
    return [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]

def case(name, cfg, n_scans, beams, az, fif=0):
    base = f"/tmp/vxdev/{name}"
    shutil.rmtree(base, ignore_errors=True)
    os.makedirs(base)
    p = vx.synth_params(n_beams=beams, n_azimuth=az)
    vx.write_sequence(p, base + "/seq", n_scans)
    t = time.time(); st_o = Oracle("port").gen_data(cfg, base + "/seq", base + "/out_port"); t_o = time.time() - t
    t = time.time(); st = vx.gen_data(cfg, base + "/seq", base + "/out_gpu", frames_in_flight=fif); t_g = time.time() - t
    bad = compare(base + "/out_port", base + "/out_gpu")
    print(f"[{name}] oracle {t_o:.2f}s ({st_o['frames']} frames)  gpu {t_g:.2f}s ({st['frames']} frames, skipped {st['skipped']})  mismatches: {bad[:8]} ({len(bad)})")
    print("   stages:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st["stages"].items()})
    return not bad

ok = True
ok &= case("small", ROOT + "/configs/test_small.cfg", 14, 16, 300)
ok &= case("small_fif1", ROOT + "/configs/test_small.cfg", 14, 16, 300, fif=1)
ok &= case("example", ROOT + "/configs/example.cfg", 12, 32, 600)
ok &= case("kitti12", ROOT + "/configs/semantic_kitti.cfg", 12, 64, 2000)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
