"""Timing of the keypoint (VIPE*) teacher on BASELINE config 4 (n = 4096 synthetic poses,
encoder (2, 1024), decoder (2, 512), 20 x 7 3-D targets); not a test.
    python tests/diag_keypoint.py [--cpu-steps K] [--graphs]  -> one JSON line (+ gpurun_out/keypoint_timing.json)
train: Keypoint_EmbeddingModel.epoch over device-resident batches (3 encoder passes, decoder,
hinge + MSE losses, backward, AdamW), device generator dropout; CUDA events around 20 steps.
apply: embed() of 65536 poses. cpu: the oracle port of the same step (torch fp32, all host
threads), bounded to K steps."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import keypoint_train_ref as T                     # noqa: E402  (cpu baseline leg only)
from vpd_b200 import keypoint                                  # noqa: E402
from vpd_b200._lib import lib                                  # noqa: E402
from vpd_b200.keypoint_train import FCPoseDecoder              # noqa: E402

N, HID, BLOCKS = 4096, 1024, 2
FWD_MFLOP = 3 * 8.534 + 2 * 0.700            # SURVEY 8(d) config 4, per sample
STEP_GFLOP = 3 * FWD_MFLOP * N / 1e3


def main():
    cpu_steps = int(sys.argv[sys.argv.index('--cpu-steps') + 1]) if '--cpu-steps' in sys.argv else 2
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    enc = keypoint.FCResNet(39, 32, BLOCKS, HID, dropout=0.2)
    dec = FCPoseDecoder(32, [512, 512], [('h36m', 140)])
    cpu_enc = {k: v.clone() for k, v in enc.state_dict().items()}
    cpu_dec = {k: v.clone() for k, v in dec.state_dict().items()}
    model = keypoint.Keypoint_EmbeddingModel(enc, {'3d': dec}, 'cuda')
    opt = model.get_optimizer(1e-4)
    model._core().use_graphs = '--graphs' in sys.argv
    batches = [{k: v.to(dev) for k, v in T.synth_batch(N, 100 + i).items()} for i in range(4)]
    model.epoch([('h36m', batches[:3])], optimizer=opt)                    # warm-up
    torch.cuda.synchronize()
    steps = 20
    l0 = lib().call('vpd_launch_count')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    contra, loss, _ = model.epoch([('h36m', [batches[i % 4] for i in range(steps)])], optimizer=opt)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3) / steps
    launches = (lib().call('vpd_launch_count') - l0) / steps
    res = {'train': {'ms_per_step': ms, 'samples_per_s': N / ms * 1e3,
                     'algorithmic_tflops': STEP_GFLOP / ms, 'launches_per_step': launches,
                     'loss': loss, 'batch': N, 'cuda_graph_replay': '--graphs' in sys.argv}}
    poses = T.synth_batch(65536, 7, with_neg=False, with_3d=False)['pose1'].to(dev)
    model.embed(poses[:4096])
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        emb = model.encoder.eval()(poses.reshape(65536, -1))
    e1.record()
    torch.cuda.synchronize()
    ms_a = e0.elapsed_time(e1) / 5
    res['apply'] = {'poses_per_s': 65536 / ms_a * 1e3, 'ms_per_65536': ms_a,
                    'algorithmic_tflops': 8.534e-6 * 65536 / ms_a * 1e3}
    # CPU leg: oracle port of the same step on the host cores
    torch.set_num_threads(os.cpu_count())
    b = T.synth_batch(N, 100)
    masks = [T.replay_masks(N, HID, BLOCKS, 3, 0.2)]
    T.zipped_step(cpu_enc, cpu_dec, dec.fcn_keys, [('h36m', b)], masks, 0.2, BLOCKS)   # warm-up
    t0 = time.perf_counter()
    for _ in range(cpu_steps):
        T.zipped_step(cpu_enc, cpu_dec, dec.fcn_keys, [('h36m', b)], masks, 0.2, BLOCKS)
    dt = (time.perf_counter() - t0) / cpu_steps
    res['cpu_baseline'] = {'samples_per_s': N / dt, 'ms_per_step': dt * 1e3, 'cores': os.cpu_count(),
                           'kind': 'port', 'sample': '{} steps of {} samples, forward + backward '
                           '(no optimizer), torch fp32'.format(cpu_steps, N)}
    print(json.dumps(res))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/keypoint_timing.json', 'w') as fp:
        json.dump(res, fp)


if __name__ == '__main__':
    main()
