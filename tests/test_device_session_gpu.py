"""Device-resident click sessions (csrc/session.cu, inference/device_session.py) against the host predictor that mirrors
the reference (inference/predictor.py + transforms.py <- isegm/inference/predictors/base.py, transforms/zoom_in.py, flip.py):
zoom-in regions and network point rows bit-exact, cropped network inputs and pasted probability maps within 2e-6."""
import numpy as np
import pytest
import torch

from pvpuformer_b200.inference.clicker import Clicker
from pvpuformer_b200.inference.evaluation import _repack_points, evaluate_lockstep, get_iou
from pvpuformer_b200.inference.predictor import vpu_eval_predictor

pytestmark = pytest.mark.gpu
T = 448


class BlobNet(torch.nn.Module):
    """Batch-independent stand-in network with masks that follow the image and react to clicks (so the zoom-in region
    moves): logits = 8 (R - 0.5) + 1.5 (prev - 0.5) + sum of +-3 Gaussian bumps at the positive / negative clicks."""
    with_prev_mask = True
    num_max_points = 24

    def forward(self, image, points, prompts=None, as_prompt_type=0):
        B, _, h, w = image.shape
        yy = torch.arange(h, device=image.device, dtype=torch.float32).view(1, 1, h, 1)
        xx = torch.arange(w, device=image.device, dtype=torch.float32).view(1, 1, 1, w)
        p = points.to(torch.float32)
        n = p.shape[1] // 2
        sign = torch.cat([torch.ones(n), -torch.ones(n)]).to(image.device).view(1, 2 * n, 1, 1)
        valid = (p[:, :, 2] >= 0).float().view(B, 2 * n, 1, 1)
        d2 = (yy - p[:, :, 0].view(B, 2 * n, 1, 1)) ** 2 + (xx - p[:, :, 1].view(B, 2 * n, 1, 1)) ** 2
        bump = (3.0 * sign * valid * torch.exp(-d2 / (2 * 30.0 ** 2))).sum(dim=1, keepdim=True)
        return {"instances": 8.0 * (image[:, 0:1] - 0.5) + 1.5 * (image[:, 3:4] - 0.5) + bump}


def _smooth(rng, H, W, cells):
    g = torch.from_numpy(rng.random((1, 1, cells, cells)).astype(np.float32))
    return torch.nn.functional.interpolate(g, size=(H, W), mode="bicubic", align_corners=True)[0, 0].numpy()


def _samples(n, H, W, seed):
    rng = np.random.default_rng(seed)
    out = []
    yy, xx = np.mgrid[:H, :W]
    for i in range(n):
        cy, cx = rng.uniform(0.25 * H, 0.75 * H), rng.uniform(0.25 * W, 0.75 * W)
        ry, rx = rng.uniform(0.08 * H, 0.3 * H), rng.uniform(0.08 * W, 0.3 * W)
        gt = ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1).astype(np.int32)
        r = np.clip(0.25 + 0.5 * gt + 0.6 * (_smooth(rng, H, W, 6) - 0.5), 0, 1)
        img = np.stack([r, _smooth(rng, H, W, 5), _smooth(rng, H, W, 4)], axis=2)
        out.append(((img * 255).astype(np.uint8), gt))
    return out


@pytest.mark.parametrize("H,W", [(448, 448), (300, 400), (600, 520)])
def test_device_sessions_track_host_predictors_click_by_click(H, W):
    from pvpuformer_b200.inference.device_session import DeviceClickSessions
    dev = torch.device("cuda:0")
    net = BlobNet().to(dev)
    samples = _samples(5, H, W, seed=H + W)
    S, K = len(samples), 7
    eng = DeviceClickSessions([s[0] for s in samples], [s[1] for s in samples], dev, max_clicks=K)
    hosts = [vpu_eval_predictor(net, dev) for _ in samples]
    clickers = [Clicker(gt_mask=s[1]) for s in samples]
    masks = [np.zeros_like(s[1]) for s in samples]
    for p, s in zip(hosts, samples):
        p.set_input_image(s[0])
    eng.clicker_step(0)
    regions = set()
    ident = exact_prev = 0
    with torch.no_grad():
        for k in range(K):
            image, points = eng.prepare()
            roi = eng.roi.cpu().numpy()
            host_logits = []
            for s in range(S):
                clickers[s].make_next_click(masks[s])
                image_nd, points_nd, _ = hosts[s].prepare_inputs(clickers[s], None, samples[s][1], 0)
                assert tuple(roi[s]) == tuple(int(v) for v in hosts[s].zoom_in._object_roi), (k, s)
                regions.add(tuple(roi[s]))
                want = _repack_points(points_nd.to(torch.float64), eng.n_half)
                got = torch.stack([points[s], points[S + s]])
                assert torch.equal(got, want), (k, s, got, want)
                got_img = torch.stack([image[s], image[S + s]])
                assert (got_img - image_nd).abs().max().item() <= 2e-6, (k, s)
                if tuple(roi[s]) == (0, T - 1, 0, T - 1) and (H, W) == (T, T):
                    assert torch.equal(got_img[:, :3], image_nd[:, :3])            # identity-sized region: bit for bit
                    exact_prev += int(torch.equal(got_img[:, 3], image_nd[:, 3]))
                    ident += 1
                host_logits.append(net(image_nd, points_nd)["instances"])
            eng.finish(net(image, points)["instances"])
            eng.clicker_step(k + 1)
            ious = eng.ious(k + 1)
            for s in range(S):
                full = hosts[s].finish_prediction(host_logits[s], (T, T))
                hosts[s].prev_prediction = full
                assert (eng.prev_probs[s] - full[0, 0]).abs().max().item() <= 2e-6, (k, s)
                masks[s] = full.cpu().numpy()[0, 0] > 0.49
                dm = eng.pred[s].cpu().numpy().astype(bool)
                assert (dm != masks[s]).sum() <= 2, (k, s, (dm != masks[s]).sum())
                assert abs(ious[s] - get_iou(samples[s][1], masks[s])) <= 1e-4
                masks[s] = dm                       # keep the two click sequences in step if a borderline pixel flipped
            for s in range(S):
                got = [(c.is_positive, c.coords) for c in eng.clicks_list(s)]
                assert got == [(bool(c.is_positive), tuple(int(v) for v in c.coords)) for c in clickers[s].clicks_list], (k, s)
    assert len(regions) > S, "the zoom-in region never moved: the test does not exercise the crop path"
    print("identity-sized regions: %d, previous-probability channel bit-identical in %d" % (ident, exact_prev))


def test_device_session_loop_matches_host_loop_with_early_stop():
    dev = torch.device("cuda:0")
    net = BlobNet().to(dev)
    samples = _samples(7, 360, 480, seed=5)
    for thr in (1.01, 0.9):
        host = evaluate_lockstep(samples, net, dev, thr, max_clicks=8, micro_batch=4)
        st = {}
        devs = evaluate_lockstep(samples, net, dev, thr, max_clicks=8, micro_batch=4, device_session=True, stats=st)
        assert [len(a) for a in host] == [len(b) for b in devs], (thr, [len(a) for a in host], [len(b) for b in devs])
        for a, b in zip(host, devs):
            assert b.dtype == np.float32 and np.abs(a - b).max() <= 1e-4, (thr, a, b)
        assert st["click_forwards"] == 2 * sum(len(a) for a in host)
    assert any(len(a) < 8 for a in host), "no session stopped early at IoU 0.9"


def test_device_session_loop_with_the_cuda_forward():
    """The real network: on 448 x 448 images at random init the regions stay identity-sized most of the time, where the
    device sessions reproduce the host loop bit for bit; elsewhere IoU agrees to 1e-4."""
    from pvpuformer_b200.config import make_config
    from pvpuformer_b200.inference.datasets import SyntheticEllipseDataset
    from pvpuformer_b200.model import build_model
    from pvpuformer_b200.weights import synthetic_state_dict
    dev = torch.device("cuda:0")
    cfg = make_config("vit_base")
    m = build_model("vit_base", state_dict=synthetic_state_dict(cfg, 0), device=dev)
    m.want_aux = False
    ds = SyntheticEllipseDataset(5, seed0=70)
    samples = [(ds.get_sample(i).image, ds.get_sample(i).gt_mask(1)) for i in range(5)]
    host = evaluate_lockstep(samples, m, dev, 1.01, max_clicks=4, micro_batch=3, device_clicker=True)
    devs = evaluate_lockstep(samples, m, dev, 1.01, max_clicks=4, micro_batch=3, device_session=True)
    for a, b in zip(host, devs):
        assert a.dtype == b.dtype and a.shape == b.shape and np.abs(a - b).max() <= 1e-4, (a, b)


def test_device_sessions_refuse_what_they_cannot_do():
    from pvpuformer_b200.inference.device_session import DeviceClickSessions
    dev = torch.device("cuda:0")
    a, b = _samples(1, 300, 400, 1)[0], _samples(1, 320, 400, 2)[0]
    with pytest.raises(ValueError):
        DeviceClickSessions([a[0], b[0]], [a[1], b[1]], dev)
    with pytest.raises(ValueError):
        DeviceClickSessions([a[0]], [a[1]], dev, max_clicks=30)


def test_session_kernels_on_large_non_square_images_and_ignore_labels():
    """1000 x 1400 images (ZoomIn crops are real down-scales there), ground truth with ignore labels: the device sessions track
    the host predictors for a few clicks."""
    from pvpuformer_b200.inference.device_session import DeviceClickSessions
    dev = torch.device("cuda:0")
    net = BlobNet().to(dev)
    samples = _samples(2, 1000, 1400, seed=77)
    rng = np.random.default_rng(3)
    samples = [(im, np.where(rng.random(gt.shape) < 0.02, -1, gt).astype(np.int32)) for im, gt in samples]
    S, K = len(samples), 4
    eng = DeviceClickSessions([s[0] for s in samples], [s[1] for s in samples], dev, max_clicks=K)
    hosts = [vpu_eval_predictor(net, dev) for _ in samples]
    clickers = [Clicker(gt_mask=s[1]) for s in samples]
    masks = [np.zeros(s[1].shape, bool) for s in samples]
    for p, s in zip(hosts, samples):
        p.set_input_image(s[0])
    eng.clicker_step(0)
    with torch.no_grad():
        for k in range(K):
            image, points = eng.prepare()
            roi = eng.roi.cpu().numpy()
            logits_h = []
            for s in range(S):
                clickers[s].make_next_click(masks[s])
                image_nd, points_nd, _ = hosts[s].prepare_inputs(clickers[s], None, samples[s][1], 0)
                assert tuple(roi[s]) == tuple(int(v) for v in hosts[s].zoom_in._object_roi), (k, s)
                assert torch.equal(torch.stack([points[s], points[S + s]]), _repack_points(points_nd.to(torch.float64), eng.n_half))
                assert (torch.stack([image[s], image[S + s]]) - image_nd).abs().max().item() <= 2e-6
                logits_h.append(net(image_nd, points_nd)["instances"])
            eng.finish(net(image, points)["instances"])
            eng.clicker_step(k + 1)
            ious = eng.ious(k + 1)
            for s in range(S):
                full = hosts[s].finish_prediction(logits_h[s], (T, T))
                hosts[s].prev_prediction = full
                assert (eng.prev_probs[s] - full[0, 0]).abs().max().item() <= 2e-6
                dm = eng.pred[s].cpu().numpy().astype(bool)
                assert (dm != (full.cpu().numpy()[0, 0] > 0.49)).sum() <= 4
                masks[s] = dm
                assert abs(ious[s] - get_iou(samples[s][1], dm)) <= 1e-12
