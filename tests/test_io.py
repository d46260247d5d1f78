"""Caller-facing plumbing (fftvis_b200/io.py): UVData array layout (CPU) and the
time-block streaming writer + hera_sim-shaped simulator (GPU)."""
import json
import types

import numpy as np
import pytest


def test_uvdata_arrays_layout_polarized_and_unpolarized():
    from fftvis_b200.io import POL_NUMS_LINEAR, fill_uvdata, uvdata_arrays
    rng = np.random.default_rng(0)
    ants = {0: np.array([0.0, 0.0, 0.0]), 3: np.array([14.6, 0.0, 0.0]), 7: np.array([7.3, 12.6, 0.5])}
    baselines = [(0, 0), (0, 3), (0, 7), (3, 7)]
    times = 2459845.0 + np.arange(3) * 1e-3
    freqs = np.linspace(100e6, 110e6, 5)
    vis = rng.normal(size=(5, 3, 2, 2, 4)) + 1j * rng.normal(size=(5, 3, 2, 2, 4))
    a = uvdata_arrays(vis, ants, baselines, times, freqs)
    assert a["data_array"].shape == (12, 5, 4) and (a["Nblts"], a["Nbls"], a["Ntimes"], a["Nfreqs"], a["Npols"]) == (12, 4, 3, 5, 4)
    assert np.array_equal(a["polarization_array"], POL_NUMS_LINEAR)
    # time-major blt axis; pols xx, yy, xy, yx from [[xx, xy], [yx, yy]]
    for t in range(3):
        for b in range(4):
            blt = t * 4 + b
            assert a["time_array"][blt] == times[t]
            assert (a["ant_1_array"][blt], a["ant_2_array"][blt]) == baselines[b]
            np.testing.assert_array_equal(a["data_array"][blt, :, 0], vis[:, t, 0, 0, b])
            np.testing.assert_array_equal(a["data_array"][blt, :, 1], vis[:, t, 1, 1, b])
            np.testing.assert_array_equal(a["data_array"][blt, :, 2], vis[:, t, 0, 1, b])
            np.testing.assert_array_equal(a["data_array"][blt, :, 3], vis[:, t, 1, 0, b])
    np.testing.assert_allclose(a["uvw_array"][3], ants[7] - ants[3])
    assert a["baseline_array"][1] == 2048 * 0 + 3 + 2 ** 16
    un = uvdata_arrays(vis[:, :, 0, 0, :], ants, baselines, times, freqs)
    assert un["data_array"].shape == (12, 5, 1) and un["polarization_array"].tolist() == [1]
    uvd = fill_uvdata(types.SimpleNamespace(), a)
    assert uvd.flag_array.shape == (12, 5, 4) and not uvd.flag_array.any() and uvd.nsample_array.min() == 1.0
    with pytest.raises(ValueError):
        uvdata_arrays(vis, ants, baselines[:3], times, freqs)


@pytest.mark.gpu
@pytest.mark.parametrize("polarized", [False, True])
def test_simulate_to_npy_streams_time_blocks(tmp_path, polarized):
    from fftvis_b200 import GaussianBeam, HERA_LOCATION, simulate_vis, synth
    from fftvis_b200.io import FFTVisB200, simulate_to_npy, uvdata_arrays
    ants = synth.hex_array(2)
    freqs = np.linspace(100e6, 120e6, 4)
    times = 2459845.0 + np.arange(7) * 30.0 / 86400.0          # 7 steps in blocks of 3: a ragged last block
    ra, dec, flux = synth.random_sky(300, freqs, seed=1)
    beam = GaussianBeam(diameter=14.0)
    kw = dict(precision=2, polarized=polarized, eps=1e-12)
    want = simulate_vis(ants, flux, ra, dec, freqs, times, beam, HERA_LOCATION, **kw)
    meta = simulate_to_npy(tmp_path / "vis.npy", ants, flux, ra, dec, freqs, times, beam, HERA_LOCATION, time_block=3, **kw)
    got = np.load(tmp_path / "vis.npy", mmap_mode="r")
    assert tuple(meta["shape"]) == want.shape == got.shape
    np.testing.assert_array_equal(np.asarray(got), want)       # same kernels, same order: bit-identical
    side = json.loads((tmp_path / "vis.npy.json").read_text())
    assert len(side["times"]) == 7 and len(side["freqs"]) == 4 and len(side["baselines"]) == want.shape[-1]
    dm = types.SimpleNamespace(ants=ants, freqs=freqs, times=times, ra=ra, dec=dec, fluxes=flux, beams=[beam],
                               telescope_loc=HERA_LOCATION)
    arrays = FFTVisB200(**kw).simulate(dm)
    ref = uvdata_arrays(want, ants, [tuple(b) for b in meta["baselines"]], times, freqs)
    np.testing.assert_array_equal(arrays["data_array"], ref["data_array"])
