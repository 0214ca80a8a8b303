"""Per-launch cost of the decoder-row kernel vs. row count (raw C-ABI calls, back-to-back on one stream).

Separates the fixed cost of a launch (prologue, pipeline fill, tail) from the per-wave cost: a "wave" is one 128-row
tile on each of the 74 CTA pairs = 9,472 rows.  Usage: python tools/launch_overhead.py [engine]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
_abi = importlib.import_module("dist-renderer_b200._abi"); planm = importlib.import_module("dist-renderer_b200.plan")
fn = importlib.import_module("dist-renderer_b200.functional")
engine = sys.argv[1] if len(sys.argv) > 1 else "tc"
dec = synth.make_decoder("B").cuda(); lat = synth.make_latent().cuda()
plan = planm.plan_for(dec)
eng = fn.resolve_engine(plan, engine)
st = torch.cuda.current_stream().cuda_stream
net, eng, keep = plan.net_for(lat, eng, st)
lib = _abi.lib()
g = torch.Generator().manual_seed(11)
N = 262144
pts = ((torch.rand(N, 3, generator=g) - 0.5) * 1.2).cuda()
sdf = torch.empty(N, device="cuda")
ndev = torch.zeros(1, dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
wave = 128 * 74
def measure(n, use_dev):
    ndev.fill_(n)
    reps = 40 if n < 100000 else 10

    def go():
        if use_dev:
            _abi.check(lib.dist_decoder_forward(net, eng, _abi.ptr(pts), N, _abi.ptr(ndev), 0.1, _abi.ptr(sdf), st))
        else:
            _abi.check(lib.dist_decoder_forward(net, eng, _abi.ptr(pts), n, None, 0.1, _abi.ptr(sdf), st))
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        go()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


# the two modes alternate (three rounds, best of) so that clock drift over the run does not read as a difference between them
print("rows (waves): host-side count | device-side count, grid sized for %d rows" % N)
for n in (1, 128, wave // 2, wave, wave + 128, 2 * wave, 3 * wave, 4 * wave, 8 * wave, 16 * wave, N):
    best = [1e30, 1e30]
    for _ in range(3):
        for k, use_dev in enumerate((False, True)):
            best[k] = min(best[k], measure(n, use_dev))
    print("  n=%7d (%5.2f waves): %8.1f us/launch | %8.1f us/launch" % (n, n / wave, best[0], best[1]))
