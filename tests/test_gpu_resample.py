"""GPU parity of the pchip resampler (seqik_pchip_resample_f32/_f64) against scipy.interpolate.pchip_interpolate called
exactly as the reference's utils.interpolate_signal calls it (seqikpy/utils.py:332-349)."""
import numpy as np
import pytest
from scipy.interpolate import pchip_interpolate

pytestmark = pytest.mark.gpu


def reference_resample(signal, original_ts, new_ts):
    total = signal.shape[0] * original_ts
    return np.array(pchip_interpolate(np.arange(0, total, original_ts), signal, np.arange(0, total, new_ts)))


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available()
    from seqikpy_b200 import engine
    return torch, engine


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-6)])
def test_resample_grooming_angles(eng, grooming_leg, dtype, tol):
    """The shipped joint angles of the grooming trial (6000 frames at 100 Hz) onto a 1 kHz grid, all 14 DOFs at once as an
    interleaved (2, 6000, 7) tensor; the samples past the last knot are the last cubic piece extrapolated, like scipy."""
    torch, engine = eng
    ang = grooming_leg["ref_angles"]                                      # (2, 6000, 7)
    out = engine.pchip_resample(torch.tensor(ang, dtype=getattr(torch, dtype), device="cuda"), 0.01, 0.001).cpu().numpy()
    assert out.shape == (2, 60000, 7)
    for leg in range(2):
        for d in range(7):
            ref = reference_resample(ang[leg, :, d], 0.01, 0.001)
            assert np.abs(out[leg, :, d] - ref).max() < tol * max(1.0, np.abs(ref).max()), (leg, d)


def test_resample_shapes_and_edge_cases(eng):
    torch, engine = eng
    rng = np.random.default_rng(3)
    for case in range(36):
        n = int(rng.integers(2, 50))
        ts, new_ts = [(0.01, 0.001), (1.0, 0.5), (0.005, 0.0007), (1 / 30, 1 / 100), (0.01, 0.025), (0.1, 0.3)][case % 6]
        y = rng.normal(size=(3, n)).cumsum(axis=1)
        if case % 3 == 0:
            y[:, n // 3:n // 2 + 1] = y[:, n // 3:n // 3 + 1]            # plateaus: zero secant slopes
        if case % 4 == 0:
            y = np.round(y)                                              # many ties and sign changes
        if len(np.arange(0, n * ts, ts)) != n:                           # numpy's arange over-ran: the reference raises there
            continue
        out = engine.pchip_resample(torch.tensor(y, dtype=torch.float64, device="cuda"), ts, new_ts).cpu().numpy()
        for r in range(3):
            ref = reference_resample(y[r], ts, new_ts)
            assert out[r].shape == ref.shape
            assert np.abs(out[r] - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()), (case, n, ts, new_ts)
    # monotone data stays monotone (the point of pchip), and two samples give a straight line
    mono = np.cumsum(rng.uniform(0, 1, size=(1, 40)), axis=1)
    out = engine.pchip_resample(torch.tensor(mono, dtype=torch.float32, device="cuda"), 0.01, 0.001).cpu().numpy()[0]
    assert np.all(np.diff(out[:391]) >= -1e-6)
    line = engine.pchip_resample(torch.tensor([[1.0, 3.0]], dtype=torch.float64, device="cuda"), 1.0, 0.25).cpu().numpy()[0]
    assert np.allclose(line, 1.0 + 2.0 * np.arange(0, 2.0, 0.25))
    # +-inf samples: zeroed together with the last sample of that series (the reference's retry); NaN raises
    y = rng.normal(size=(2, 30))
    y[0, 7] = np.inf
    fixed = y.copy()
    fixed[0, 7] = 0
    fixed[0, -1] = 0
    out = engine.pchip_resample(torch.tensor(y, dtype=torch.float64, device="cuda"), 0.01, 0.002).cpu().numpy()
    for r in range(2):
        assert np.abs(out[r] - reference_resample(fixed[r], 0.01, 0.002)).max() < 1e-11
    y[1, 3] = np.nan
    with pytest.raises(ValueError):
        engine.pchip_resample(torch.tensor(y, dtype=torch.float64, device="cuda"), 0.01, 0.002)
    with pytest.raises(ValueError):
        engine.pchip_resample(torch.zeros((2, 1), dtype=torch.float64, device="cuda"), 0.01, 0.002)


def test_utils_interpolate_like_the_reference(eng, grooming_leg):
    """utils.interpolate_signal / interpolate_joint_angles (reference utils.py:332-359) on the device kernel: same keys,
    lengths and values as the reference's scipy call."""
    from seqikpy_b200.utils import interpolate_joint_angles, interpolate_signal
    keys = [str(k) for k in grooming_leg["angle_keys"]]
    ang = {k: grooming_leg["ref_angles"].reshape(-1, 7)[:6000, i % 7].copy() if i < 7 else grooming_leg["ref_angles"][1][:, i - 7].copy()
           for i, k in enumerate(keys)}
    ang["short"] = np.cos(np.arange(0, 3.0, 0.01))                       # a second length in the same dictionary
    out = interpolate_joint_angles(ang, original_ts=0.01, new_ts=0.001)
    assert list(out.keys()) == list(ang.keys())
    for k, v in ang.items():
        ref = reference_resample(v, 0.01, 0.001)
        assert out[k].shape == ref.shape and out[k].dtype == np.float64
        assert np.abs(out[k] - ref).max() < 1e-12 * max(1.0, np.abs(ref).max()), k
    assert np.allclose(interpolate_signal(np.arange(5.0), 1.0, 0.5)[:8], np.arange(0, 4.0, 0.5))
    two = np.stack([np.sin(np.arange(50) * 0.3), np.arange(50) ** 1.5], axis=1)             # (n, k) along axis 0
    got = interpolate_signal(two, 0.02, 0.005)
    assert got.shape == (200, 2)
    for j in range(2):
        assert np.abs(got[:, j] - reference_resample(two[:, j], 0.02, 0.005)).max() < 1e-10
