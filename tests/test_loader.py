"""Frame reader / loader tests.  CPU: record files, shuffle-buffer semantics.  GPU: the pinned async
loader end to end against a numpy restatement of analyzer.py:111-127."""
import os

import numpy as np
import pytest
import torch

import analyzer
from oracle import convvae_ref as R


def _write_bins(tmp_path, n_files=3, frames=(40, 25, 61), seed=0):
    rs = np.random.RandomState(seed)
    paths, allrec = [], []
    for i in range(n_files):
        rec = rs.randn(frames[i], analyzer.FEAT_DIM).astype(np.float32)
        rec[:, -1] = i % 10                                    # speaker id as float (analyzer.py:62-66)
        p = tmp_path / ("spk%d" % i); p.mkdir()
        f = p / "utt.bin"; rec.tofile(str(f)); paths.append(str(f)); allrec.append(rec)
    return paths, np.concatenate(allrec)


def test_record_constants():
    assert (analyzer.SP_DIM, analyzer.FEAT_DIM, analyzer.RECORD_BYTES) == (513, 1029, 4116)
    assert len(analyzer.SPEAKERS) == 10


def test_shuffle_reader_is_a_permutation_stream(tmp_path):
    paths, allrec = _write_bins(tmp_path)
    rf = analyzer.RecordFiles(paths)
    assert rf.n_records() == len(allrec)
    rd = analyzer.ShuffleReader(rf, batch_size=7, capacity=32, min_after_dequeue=16, seed=1)
    keys = {tuple(r[:4]) for r in allrec}
    seen = []
    out = np.empty((7, analyzer.FEAT_DIM), np.float32)
    for _ in range(3 * len(allrec) // 7):
        rd.next_batch(out)
        assert rd.fill >= rd.min_after                         # never below min_after_dequeue
        for r in out:
            assert tuple(r[:4]) in keys                        # every row is a real record, unmodified
            seen.append(tuple(r[:4]))
    # sampling without replacement from the buffer: each record appears about once per epoch
    counts = {k: seen.count(k) for k in keys}
    assert max(counts.values()) <= 4 and sum(counts.values()) == len(seen)
    assert len(set(seen[:7])) == 7                             # no duplicates inside a batch
    assert seen[:7] != [tuple(r[:4]) for r in allrec[:7]]      # and it is shuffled


def test_whole_file_reader(tmp_path):
    paths, allrec = _write_bins(tmp_path, 1, (33,))
    feats = list(analyzer.read_whole_features(os.path.join(str(tmp_path), "*", "*.bin")))
    assert len(feats) == 1 and feats[0]["sp"].shape == (33, 513) and feats[0]["speaker"].dtype == np.int64
    assert np.array_equal(feats[0]["f0"], allrec[:, 1026]) and np.array_equal(feats[0]["en"], allrec[:, 1027])


@pytest.mark.gpu
def test_async_loader_matches_reference_semantics(tmp_path, arch):
    from vae_npvc_b200.engine import Engine
    paths, allrec = _write_bins(tmp_path)
    rs = np.random.RandomState(3)
    xmin = -2.5 - rs.rand(513); xmax = 2.5 + rs.rand(513)        # a few percent of the N(0,1) data clips
    eng = Engine(arch)
    norm = analyzer.Tanhize(xmin=xmin, xmax=xmax, engine=eng)
    image, label = analyzer.read(os.path.join(str(tmp_path), "*", "*.bin"), batch_size=16, capacity=64,
                                 min_after_dequeue=32, normalizer=norm, engine=eng)
    lut = {tuple(np.round(R.tanhize_forward(r[:513].astype(np.float64), xmin, xmax)[:6], 5)): r for r in allrec}
    for _ in range(5):
        x, y = image.dequeue()
        assert x.shape == (16, 1, 513, 1) and y.shape == (16,) and y.dtype == torch.int64 and x.is_cuda
        xs = x.reshape(16, 513).cpu().numpy()
        for i in range(16):
            r = lut[tuple(np.round(xs[i, :6].astype(np.float64), 5))]
            err = np.abs(xs[i] - R.tanhize_forward(r[:513].astype(np.float64), xmin, xmax))
            assert err.max() < 1e-5, (float(err.max()), int(err.argmax()))
            assert int(y[i]) == int(r[-1])
    xp, yp = image.dequeue(peek=True)
    xq, yq = label.dequeue()
    assert torch.equal(xp, xq) and torch.equal(yp, yq)          # both handles share one queue; peek does not consume
    image.loader.close()


@pytest.mark.gpu
def test_main_and_convert_drivers_end_to_end(tmp_path, arch, monkeypatch):
    """README command lines of the reference, on synthetic VCC2016-shaped data:
    python main.py --model ConvVAE --trainer VAETrainer --architecture <json>   (main.py:45-74)
    python convert.py --src SF1 --trg TM3 --model ConvVAE --checkpoint <ckpt>  (convert.py:66-116)"""
    import json
    import main as main_driver
    import convert as convert_driver
    monkeypatch.chdir(tmp_path)
    rs = np.random.RandomState(0)
    (tmp_path / "etc").mkdir()
    xmin = (-3 - rs.rand(513)); xmax = (3 + rs.rand(513))
    xmin.astype(np.float64).tofile("etc/xmin.npf"); xmax.astype(np.float64).tofile("etc/xmax.npf")   # float64 (SURVEY 5 dtype trap)
    for spk in ("SF1", "TM3"):
        np.array([5.0, 0.3], np.float32).tofile("etc/%s.npf" % spk)
    for split in ("Training Set", "Testing Set"):
        for si, spk in enumerate(("SF1", "TM3")):
            d = tmp_path / "dataset" / "vcc2016" / "bin" / split / spk; d.mkdir(parents=True)
            rec = rs.randn(90, analyzer.FEAT_DIM).astype(np.float32)
            rec[:, 1026] = np.abs(rec[:, 1026]) * 100 + 80          # f0 > 1
            rec[:, -1] = analyzer.SPEAKERS.index(spk)
            rec.tofile(str(d / "100001.bin"))
    a = dict(arch); a["training"] = dict(arch["training"], max_iter=12, batch_size=32)
    json.dump(a, open("architecture-vae-test.json", "w"))
    main_driver.main(["--model", "ConvVAE", "--trainer", "VAETrainer", "--architecture", "architecture-vae-test.json"])
    runs = sorted((tmp_path / "logdir" / "train").iterdir())
    assert len(runs) == 1 and (runs[0] / "architecture-vae-test.json").exists() and (runs[0] / "training.log").exists()
    ckpt = runs[0] / "model.ckpt-12"
    assert ckpt.exists()
    convert_driver.main(["--src", "SF1", "--trg", "TM3", "--model", "ConvVAE", "--checkpoint", str(ckpt)])
    outs = list((tmp_path / "logdir" / "output").glob("*/SF1-TM3-100001.bin"))
    assert len(outs) == 1
    conv = np.fromfile(str(outs[0]), np.float32).reshape(-1, analyzer.FEAT_DIM)
    assert conv.shape == (90, analyzer.FEAT_DIM) and np.isfinite(conv).all() and (conv[:, -1] == analyzer.SPEAKERS.index("TM3")).all()
    src = np.fromfile(str(tmp_path / 'dataset' / 'vcc2016' / 'bin' / 'Testing Set' / 'SF1' / '100001.bin'), np.float32).reshape(-1, analyzer.FEAT_DIM)
    assert np.array_equal(conv[:, 513:1026], src[:, 513:1026]) and np.array_equal(conv[:, 1027], src[:, 1027])     # ap, en pass through


def test_trainer_writes_reference_summary_tags(tmp_path):
    """trainer/vae.py status + summaries without a GPU: the training.log line format of the reference
    (trainer/vae.py:31-52) and the TensorBoard scalar tags of model/vae.py:132-133."""
    import importlib
    T = importlib.import_module("trainer.vae").VAETrainer
    arch = {"training": {"lr": 1e-4, "beta1": 0.5, "beta2": 0.999, "max_iter": 1}}
    dirs = {"logdir": str(tmp_path / "train")}
    tr = T({"G": 0.0}, arch, None, dirs)
    tr.global_step = 7
    msg = tr._refresh_status()
    assert msg.startswith("Iter 00007: log P(x|z, y) = ") and "D_KL(z) = " in msg
    assert "Iter 00007" in open(os.path.join(dirs["logdir"], "training.log")).read()
    if tr._write_summaries([1.0, 2.5, -3.5]):
        from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
        acc = EventAccumulator(dirs["logdir"]); acc.Reload()
        assert set(acc.Tags()["scalars"]) == {"KL-div", "logPx"}
        assert acc.Scalars("KL-div")[0].value == 2.5 and acc.Scalars("logPx")[0].step == 7


def test_trainer_checkpoint_resume_roundtrip(tmp_path):
    """save -> restore (trainer/vae.py:78-84 Supervisor autosave / restore-on-start, util/wrapper.py:32-62 `load`):
    the variables under their TF names, the Adam slots and global_step come back; the newest
    ``model.ckpt-<step>`` of the logdir is the one taken; shape mismatches are refused.  No GPU needed: the
    machine is a stand-in exposing what the trainer touches (theta, variables())."""
    import importlib
    import torch
    tv = importlib.import_module("trainer.vae")

    class Machine(object):
        def __init__(self, seed):
            g = torch.Generator().manual_seed(seed)
            self.theta = torch.randn(10, generator=g)
            self.state = torch.zeros(4, dtype=torch.int64)       # npvc_step_state stand-in: {seed, draws, step, -}

        def variables(self):
            return {"Encoder/dense/kernel": self.theta[:6].view(2, 3), "Encoder/dense/bias": self.theta[6:]}

    arch = {"training": {"lr": 1e-4, "beta1": 0.5, "beta2": 0.999, "max_iter": 1}}
    dirs = {"logdir": str(tmp_path / "train"), "restore_from": str(tmp_path / "train")}
    assert tv.latest_checkpoint(dirs["logdir"]) is None
    m1 = Machine(1)
    t1 = tv.VAETrainer({"G": 0.0}, arch, None, dirs); t1.machine = m1
    assert t1.restore() is None                                   # nothing to restore yet
    st = t1._ensure_state(m1)
    st["m"].fill_(0.25); st["v"].fill_(0.5)
    t1.global_step = 9; t1.save()
    t1.global_step = 120; m1.theta.mul_(2.0); st["m"].fill_(0.75); t1.save()
    assert tv.latest_checkpoint(dirs["logdir"]).endswith("model.ckpt-120")      # numeric, not lexicographic, order

    m2 = Machine(2)
    t2 = tv.VAETrainer({"G": 0.0}, arch, None, dirs); t2.machine = m2
    assert t2.restore() == 120 and t2.global_step == 120
    assert m2.state.tolist()[1:3] == [120, 120]                   # the device-side counters follow the restored step
    assert torch.equal(m2.theta, m1.theta) and torch.equal(t2._state["m"], st["m"]) and torch.equal(t2._state["v"], st["v"])
    m3 = Machine(3)
    t3 = tv.VAETrainer({"G": 0.0}, arch, None, dirs); t3.machine = m3
    assert t3.restore(ckpt="model.ckpt-9") == 9 and float(t3._state["m"][0]) == 0.25
    assert torch.allclose(m3.theta * 2.0, m1.theta)

    class Other(Machine):
        def variables(self):
            return {"Encoder/dense/kernel": self.theta[:6].view(3, 2), "Encoder/dense/bias": self.theta[6:]}
    t4 = tv.VAETrainer({"G": 0.0}, arch, None, dirs); t4.machine = Other(4)
    with pytest.raises(ValueError):
        t4.restore()


def test_convert_f0_follows_the_reference_chain(tmp_path):
    """convert.py:51-57: three chained tf.where's, each on the UPDATED value -- log where f0 > 1, the Gaussian
    transform where the log is > 1, exp where the transformed value is > 1 (so an f0 in (1, e] comes out as its
    logarithm, as in the reference); unvoiced frames (f0 = 0) pass through."""
    import convert
    etc = tmp_path / "etc"; etc.mkdir()
    np.array([5.0, 0.3], np.float32).tofile(str(etc / "SF1.npf")); np.array([4.5, 0.2], np.float32).tofile(str(etc / "TM3.npf"))
    f0 = np.array([0.0, 0.5, 2.0, 100.0, 220.0], np.float32)
    out = convert.convert_f0(f0, "SF1", "TM3", etc=str(etc))
    lf = np.log(f0[3:].astype(np.float64))
    want = np.exp((lf - 5.0) / np.float32(0.3) * np.float32(0.2) + 4.5)
    assert out.dtype == np.float32 and out[0] == 0.0 and out[1] == 0.5
    assert abs(out[2] - np.log(2.0)) < 1e-6                       # log(2) = 0.69 is not > 1: stays the logarithm
    assert np.allclose(out[3:], want, rtol=1e-5)
