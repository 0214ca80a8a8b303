"""Do decoder-row launches on different CUDA streams overlap?  (raw C-ABI calls; design input for multi-view batching)

case 1: 40 one-wave launches (16 tiles each) on one stream vs. 20 + 20 on two streams
case 2: 40 small launches + 8 launches of 65,536 rows: one stream vs. small on A, big on B"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("dist-renderer_b200"); synth = importlib.import_module("dist-renderer_b200.synth")
_abi = importlib.import_module("dist-renderer_b200._abi"); planm = importlib.import_module("dist-renderer_b200.plan")
fn = importlib.import_module("dist-renderer_b200.functional")
dec = synth.make_decoder("B").cuda(); lat = synth.make_latent().cuda()
plan = planm.plan_for(dec)
main = torch.cuda.current_stream()
net, eng, keep = plan.net_for(lat, fn.resolve_engine(plan, "tc"), main.cuda_stream)
lib = _abi.lib()
N = 65536
pts = ((torch.rand(N, 3) - 0.5) * 1.2).cuda()
outs = [torch.empty(N, device="cuda") for _ in range(2)]
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()


def launch(stream, n, out):
    _abi.check(lib.dist_decoder_forward(net, eng, _abi.ptr(pts), n, None, 0.1, _abi.ptr(out), stream.cuda_stream))


def timed(fn_):
    for _ in range(2):
        fn_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    sA.wait_stream(main); sB.wait_stream(main)
    fn_()
    main.wait_stream(sA); main.wait_stream(sB)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


small = 2048
def c1_serial():
    for _ in range(40): launch(sA, small, outs[0])
def c1_two():
    for _ in range(20):
        launch(sA, small, outs[0]); launch(sB, small, outs[1])
def c2_serial():
    for i in range(8):
        launch(sA, N, outs[1])
        for _ in range(5): launch(sA, small, outs[0])
def c2_two():
    for i in range(8):
        launch(sB, N, outs[1])
        for _ in range(5): launch(sA, small, outs[0])
for name, f in (("case1 one stream", c1_serial), ("case1 two streams", c1_two), ("case2 one stream", c2_serial), ("case2 two streams", c2_two)):
    print("%-20s %9.1f us" % (name, timed(f)), flush=True)
