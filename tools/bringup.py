"""GPU bring-up diagnostics (not a test): prints per-op and per-layer errors of the CUDA path vs the oracle."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
import numpy as np, torch
from fav_b200 import _lib, models_video, synth, utils, preprocess, consistencyChecker, stn
from oracle import net_oracle, pyoracle

dev = torch.device("cuda:0")
def T(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def section(name): print(f"\n=== {name} ===", flush=True)

def front(H, W):
    section(f"front end {H}x{W}")
    img = synth.make_frame(H, W, 1); img2 = synth.make_frame(H, W, 2)
    bw = synth.make_backward_flow(H, W, 2); fw = synth.make_forward_flow(H, W, 2)
    flow = synth.checker_to_lua(bw)
    o = pyoracle.warp_bdhw(img, flow); g = utils.warp_image(T(img), T(flow)).cpu().numpy()
    print("warp per-tap      max|d| %.3e  exact=%s" % (np.abs(o - g).max(), np.array_equal(o, g)))
    o = pyoracle.image_warp_pad(img, flow); g = utils.warp_image(T(img), T(flow), "torch.FloatTensor").cpu().numpy()
    print("warp pad-pixel    max|d| %.3e  exact=%s" % (np.abs(o - g).max(), np.array_equal(o, g)))
    sf = synth.stress_flow(H, W)
    o = pyoracle.warp_bdhw(img, sf); g = utils.warp_image(T(img), T(sf)).cpu().numpy()
    print("warp stress flow  max|d| %.3e  exact=%s" % (np.abs(o - g).max(), np.array_equal(o, g)))
    rel_o = pyoracle.consistency(bw, fw); rel_g = consistencyChecker.check(T(bw), T(fw)).cpu().numpy()
    print("consistency 3-arg mismatches %d / %d (zeros %d)" % ((rel_o != rel_g).sum(), rel_o.size, (rel_o == 0).sum()))
    img255 = np.clip(np.rint(img2 * 255), 0, 255).astype(np.float32)
    fwn = (fw + np.random.default_rng(0).normal(0, .6, fw.shape)).astype(np.float32)
    img255[:, :, : W // 2] = 128.0
    rel_o = pyoracle.consistency(bw, fwn, img255); rel_g = consistencyChecker.check(T(bw), T(fwn), T(img255)).cpu().numpy()
    print("consistency 4-arg mismatches %d / %d (zeros %d)" % ((rel_o != rel_g).sum(), rel_o.size, (rel_o == 0).sum()))
    co = pyoracle.compute_corners(img255); cg, avg = consistencyChecker.compute_corners(T(img255))
    print("corners max|d| %.3e exact=%s avg %.9g vs %.9g" % (np.abs(co - cg.cpu().numpy()).max(), np.array_equal(co, cg.cpu().numpy()), float(avg.item()), float(pyoracle.lib().orc_avg(co.ctypes.data_as(__import__('ctypes').POINTER(__import__('ctypes').c_float)), __import__('ctypes').c_int64(co.size)))))
    cert = rel_o.astype(np.float32) / 255
    o = pyoracle.min_filter(cert, 7); g = utils.min_filter(T(cert), 7).cpu().numpy()
    print("min_filter        exact=%s" % np.array_equal(o, g))
    rnd = np.random.default_rng(1).uniform(0, 1, (H, W)).astype(np.float32)
    o = pyoracle.min_filter(rnd, 7); g = utils.min_filter(T(rnd), 7).cpu().numpy()
    print("min_filter random exact=%s" % np.array_equal(o, g))
    o = pyoracle.vgg_preprocess(img); g = preprocess.vgg.preprocess(T(img)[None]).cpu().numpy()[0]
    print("preprocess        exact=%s" % np.array_equal(o, g))
    o2 = pyoracle.vgg_deprocess(o); g2 = preprocess.vgg.deprocess(T(o)[None]).cpu().numpy()[0]
    print("deprocess         exact=%s" % np.array_equal(o2, g2))
    cm = pyoracle.min_filter(cert, 7)
    o = pyoracle.temporal_input(img2, img, flow, cm)
    out7 = torch.empty((7, H, W), device=dev)
    _lib.check(_lib.lib.fav_temporal_input(_lib.dptr(T(img2)), _lib.dptr(T(img)), _lib.dptr(T(flow)), _lib.dptr(T(cm)), None, None, _lib.dptr(out7), H, W, 0, _lib.stream_ptr()))
    print("temporal_input    max|d| %.3e exact=%s" % (np.abs(o - out7.cpu().numpy()).max(), np.array_equal(o, out7.cpu().numpy())))

def net_layers(H, W, impl):
    section(f"net {impl} {H}x{W}: per-layer error vs fp64 oracle")
    net = models_video.synthetic_model("candy"); net.set_conv_impl(impl)
    ora = net_oracle.NetOracle(style="candy", dtype=torch.float64)
    x7 = pyoracle.first_frame_input(synth.make_frame(H, W, 1))
    x7[3:6] = np.random.default_rng(2).normal(0, 50, (3, H, W)); x7[6] = np.random.default_rng(3).uniform(0, 1, (H, W))
    taps = {}
    ref = ora.forward(torch.from_numpy(x7)[None], taps)[0].numpy()
    t = time.time(); out = net.forward(T(x7)[None]); torch.cuda.synchronize(); dt = time.time() - t
    for i in range(len(ora.specs) - 1):
        try:
            lo = net.layer_output(i).cpu().numpy(); r = taps[f"l{i}"][0].numpy()
            print("layer %2d %-8s shape %-16s max|d| %.3e  (max|ref| %.2f)" % (i, ora.specs[i]["kind"], str(lo.shape), np.abs(lo - r).max(), np.abs(r).max()))
        except Exception as e:
            print("layer", i, "ERR", e)
    o = out.cpu().numpy()[0]
    print("net out max|d| %.3e (max|ref| %.2f) -> /255 = %.3e   [%.1f ms incl. plan]" % (np.abs(o - ref).max(), np.abs(ref).max(), np.abs(o - ref).max() / 255, dt * 1e3))
    return net

def timing(net, H, W, n=5):
    x = torch.randn(1, 7, H, W, device=dev) * 50
    net.forward(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): net.forward(x)
    e1.record(); torch.cuda.synchronize()
    print("forward %dx%d: %.3f ms/frame" % (H, W, e0.elapsed_time(e1) / n), flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "launches", _lib.lib.fav_launch_count())
    steps = sys.argv[1:] or ["front", "simt", "tc", "time"]
    for s in steps:
        try:
            if s == "front": front(64, 96); front(100, 76); front(97, 75)
            elif s == "simt": net_layers(64, 96, "simt")
            elif s == "tc":
                net = net_layers(64, 96, "tcgen05"); net_layers(256, 256, "tcgen05")
            elif s == "time":
                net = models_video.synthetic_model("candy")
                for impl in ("tcgen05", "simt"):
                    net.set_conv_impl(impl); print(impl); timing(net, 256, 256); timing(net, 720, 1280, 3)
        except Exception:
            traceback.print_exc()
            if "CUDA" in traceback.format_exc() or "status 2" in traceback.format_exc(): 
                print("CUDA error: aborting remaining steps"); break
    print("launches", _lib.lib.fav_launch_count())
