"""TEST INFRASTRUCTURE: drives the reference's UNMODIFIED `python -m mask_cyclegan_vc.train` and
`python -m mask_cyclegan_vc.test` (train.py:175-375, test.py:85-119, saver/model_saver.py:46-123)

  arm "reference": reference model.py (oracle/_ref), CPU fp32 for training (the oracle run);
  arm "engine":    the same drivers with this repo's shim first on PYTHONPATH, so that
                   `from mask_cyclegan_vc.model import Generator, Discriminator` resolves to the
                   sm_100a engine; everything else (args, dataset, logger, saver) is the reference's.

and compares the loss trajectories and the converted utterances.  Needs oracle/_ref (built by
oracle/build_ref.sh where /root/reference exists; travels to the GPU box with the snapshot).

    python oracle/refharness/run_dropin.py <workdir> [--epochs 2] [--batch 2] [--precision c8]
"""
import argparse
import glob
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
SHIM = os.path.join(ROOT, "maskcyclegan-vc_b200", "shim")
sys.path.insert(0, HERE)


def have_ref():
    return os.path.exists(os.path.join(REF, "mask_cyclegan_vc", "train.py"))


def _env(arm, scalar_log, extra=None):
    env = dict(os.environ)
    path = [HERE, REF]
    if arm == "engine":
        path.insert(0, SHIM)
    env["PYTHONPATH"] = os.pathsep.join(path)
    env["MCGVC_SCALAR_LOG"] = scalar_log
    env["MCGVC_REF_HARNESS"] = "1"
    env.pop("CUDA_VISIBLE_DEVICES", None)
    env.update(extra or {})
    return env


def run_train(arm, work, data, epochs, batch, gpu, extra_env=None, timeout=1500):
    log = os.path.join(work, arm + "_scalars.jsonl")
    if os.path.exists(log):
        os.unlink(log)
    cmd = [sys.executable, "-m", "mask_cyclegan_vc.train", "--name", arm, "--seed", "0", "--save_dir", work,
           "--preprocessed_data_dir", data, "--speaker_A_id", "SPKA", "--speaker_B_id", "SPKB",
           "--num_epochs", str(epochs), "--batch_size", str(batch), "--steps_per_print", "1",
           "--epochs_per_save", "1", "--epochs_per_plot", "1", "--num_frames", "64", "--max_mask_len", "25",
           "--gpu_ids", gpu]
    r = subprocess.run(cmd, cwd=work, env=_env(arm, log, extra_env), capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s train.py failed (rc %d)\n%s\n%s" % (arm, r.returncode, r.stdout[-3000:], r.stderr[-3000:]))
    out = {"g_loss": [], "d_loss": []}
    with open(log) as f:
        for line in f:
            rec = json.loads(line)
            key = rec["name"].replace("/", "_")       # the logger turns g_loss into g/loss (base_logger.py:78)
            if key in out:
                out[key].append(rec["value"])
    return out, r.stdout + r.stderr


def run_test(arm, work, data, ckpt_dir, load_epoch, model_name, gpu, name, extra_env=None, timeout=900):
    log = os.path.join(work, name + "_scalars.jsonl")
    cmd = [sys.executable, "-m", "mask_cyclegan_vc.test", "--name", name, "--save_dir", work,
           "--preprocessed_data_dir", data, "--speaker_A_id", "SPKA", "--speaker_B_id", "SPKB",
           "--ckpt_dir", ckpt_dir, "--load_epoch", str(load_epoch), "--model_name", model_name, "--gpu_ids", gpu]
    r = subprocess.run(cmd, cwd=work, env=_env(arm, log, extra_env), capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s test.py failed (rc %d)\n%s\n%s" % (arm, r.returncode, r.stdout[-3000:], r.stderr[-3000:]))
    wavs = {}
    for p in sorted(glob.glob(os.path.join(work, name, "converted_audio", "*-converted_*.npy"))):
        wavs[os.path.basename(p)] = np.load(p)
    return wavs, r.stdout + r.stderr


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("work")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--precision", default=None, help="MCGVC_PRECISION for the engine arm (default: library default)")
    ap.add_argument("--ref-train-device", default="-1", help="--gpu_ids of the reference train arm (-1 = CPU fp32)")
    args = ap.parse_args()
    if not have_ref():
        raise SystemExit("oracle/_ref missing: run oracle/build_ref.sh where /root/reference exists")
    import make_synth_data
    work = os.path.abspath(args.work)
    os.makedirs(work, exist_ok=True)
    data = make_synth_data.make(os.path.join(work, "data"))
    eng_env = {"MCGVC_PRECISION": args.precision} if args.precision else {}
    no_tf32 = {"NVIDIA_TF32_OVERRIDE": "0"}      # the reference's own CUDA path defaults to TF32 convolutions

    report = {"epochs": args.epochs, "batch": args.batch, "precision": args.precision or "default"}
    ref_tr, _ = run_train("reference", work, data, args.epochs, args.batch, args.ref_train_device, no_tf32)
    eng_tr, eng_log = run_train("engine", work, data, args.epochs, args.batch, "0", eng_env)
    report["steps"] = len(ref_tr["g_loss"])
    report["reference_losses"] = ref_tr
    report["engine_losses"] = eng_tr
    report["loss_rel_dev"] = {k: [abs(e - r) / max(abs(r), 1e-12) for e, r in zip(eng_tr[k], ref_tr[k])] for k in ref_tr}
    report["engine_checkpoints"] = sorted(os.path.basename(p) for p in glob.glob(os.path.join(work, "engine", "ckpts", "*.pth.tar")))
    report["reference_checkpoints"] = sorted(os.path.basename(p) for p in glob.glob(os.path.join(work, "reference", "ckpts", "*.pth.tar")))

    # test.py on the reference-written checkpoint: engine vs the reference's own CUDA modules (TF32 off)
    ref_ck = os.path.join(work, "reference", "ckpts")
    eng_ck = os.path.join(work, "engine", "ckpts")
    ep = args.epochs
    w_eng, _ = run_test("engine", work, data, ref_ck, ep, "generator_A2B", "0", "test_engine_on_refckpt", eng_env)
    w_ref, _ = run_test("reference", work, data, ref_ck, ep, "generator_A2B", "0", "test_reference_on_refckpt", no_tf32)
    report["test_utterances"] = len(w_ref)
    report["test_wav_rel_err"] = {k: rel(w_eng[k], w_ref[k]) for k in w_ref}
    report["test_wav_frames"] = {k: int(v.shape[-1] // 80) for k, v in w_ref.items()}
    # wire compatibility the other way: the reference's modules load the engine-written checkpoint (strict)
    w_ref2, _ = run_test("reference", work, data, eng_ck, ep, "generator_B2A", "0", "test_reference_on_engckpt", no_tf32)
    w_eng2, _ = run_test("engine", work, data, eng_ck, ep, "generator_B2A", "0", "test_engine_on_engckpt", eng_env)
    report["test_wav_rel_err_engine_ckpt"] = {k: rel(w_eng2[k], w_ref2[k]) for k in w_ref2}
    print(json.dumps(report))


if __name__ == "__main__":
    main()
