import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdf_viewer_b200 as S
from tools.configs_run import timed, BB
side = 512
workload = sys.argv[1] if len(sys.argv) > 1 else "demo"
with S.SDFViewer.from_bb(BB, side, 2) as v:
    stream = torch.cuda.ExternalStream(v.stream)
    v.set_tape(S.tape.demo_tape() if workload == "demo" else S.tape.csg_tape())
    v.set_option("fill_program", 3)
    for vpt in (2, 4, 8):
        v.set_option("fill_voxels_per_thread", vpt)
        ms = timed(v, stream, v.fill_all, 10)
        print(f"{workload} MINB={os.environ.get('SDFGPU_JIT_MINB')} vpt={vpt} ctas/SM={v.get_info('last_fill_ctas_per_sm')}: {ms:.4f} ms {side**3*32/ms/1e6:.0f} GB/s", flush=True)
