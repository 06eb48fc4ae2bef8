"""GPU: the multi-device front end (drfe_pool_*): a batch cut into contiguous blocks over several
workers gives, frame by frame, exactly what one handle pair gives — with two workers on one GPU (always testable)
and with one worker per GPU when the box has several."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MC = float(np.float32(np.cos(np.pi / 12)))


@pytest.fixture(scope="module")
def frames(drfe):
    data = [drfe.synth_frame(640, 480, i % 3, 20261200 + i) for i in range(13)]
    return np.stack([d[0] for d in data]), np.stack([d[1] for d in data]), data[0][2]


@pytest.fixture(scope="module")
def single(drfe, frames):
    gray, depth, K = frames
    n = len(gray)
    orb = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=n)
    cape = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=n)
    orb.enqueue(gray); cape.enqueue_depth(depth, *K)
    return orb.download(), cape.download()


def check(out, single, n):
    (rk, rd, rc), (rseg, rpl, rnp) = single
    assert np.array_equal(out["counts"][:n], rc[:n]) and np.array_equal(out["nplanes"][:n], rnp[:n])
    assert np.array_equal(out["seg"][:n], rseg[:n])
    for f in range(n):
        c = rc[f]
        assert out["kps"][f, :c].tobytes() == rk[f, :c].tobytes() and np.array_equal(out["desc"][f, :c], rd[f, :c])
        for name in ("normal", "d", "nr_pts", "MSE", "score"):
            assert np.array_equal(out["planes"][f, :rnp[f]][name], rpl[f, :rnp[f]][name]), name


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_pool_on_one_gpu_matches_single_handles(drfe, frames, single, devices):
    gray, depth, K = frames
    pool = drfe.Pool(devices, width=640, height=480, min_cos=MC, max_merge_dist=50.0, max_batch=len(gray))
    assert pool.ndev == len(devices)
    out = pool.extract_batch(gray, depth, *K)
    check(out, single, len(gray))
    # a shorter batch through the same pool, raw 16-bit depth scaled on the device (Frame.cc:113-115)
    q = np.rint(depth * 5000).astype(np.uint16)
    out = pool.extract_batch(gray[:5], q[:5], *K, depth_factor=float(np.float32(1.0 / 5000.0)))
    check(out, single, 5)
    assert (pool.device_times()[: min(len(devices), 5)] > 0).all()
    pool.close()


def test_pool_with_pinned_buffers_and_every_gpu(drfe, frames, single):
    gray, depth, K = frames
    ndev = drfe.device_count()
    pool = drfe.Pool(list(range(ndev)), width=640, height=480, min_cos=MC, max_merge_dist=50.0, max_batch=len(gray))
    hg = drfe.host_array(gray.shape, np.uint8); hg[...] = gray
    hd = drfe.host_array(depth.shape, np.float32, write_combined=True); hd[...] = depth
    out = pool.extract_batch(hg, hd, *K)
    check(out, single, len(gray))
    pool.close()


def test_pool_argument_errors(drfe, frames):
    gray, depth, K = frames
    with pytest.raises(drfe.DrfeError):
        drfe.Pool([drfe.device_count() + 3], max_batch=4)                       # no such device
    pool = drfe.Pool([0], width=640, height=480, max_batch=4)
    with pytest.raises(drfe.DrfeError) as e:
        pool.extract_batch(gray[:5], depth[:5], *K)                             # more frames than max_batch
    assert e.value.code == drfe.ERR_ARG
    small = {"kps": np.empty((2, 10), drfe.KP_DTYPE), "desc": np.empty((2, 10, 32), np.uint8), "counts": np.empty(2, np.int32),
             "seg": np.empty((2, 480, 640), np.uint8), "planes": np.empty((2, 64), drfe.PLANE_DTYPE), "nplanes": np.empty(2, np.int32)}
    with pytest.raises(drfe.DrfeError) as e:
        pool.extract_batch(gray[:2], depth[:2], *K, out=small)                   # 10 keypoint slots per frame
    assert e.value.code == drfe.ERR_CAPACITY and "device 0" in str(e.value)
    pool.close()
