"""One RAGGED run through the batched segment driver (xiaoicesing_io_b200.segments): N synthetic segments with lengths drawn
uniformly from [min_s, max_s] seconds, bucketed into ragged batches, DDIM-20 (config 1's sampler) on a WaveNet 20x256.  Prints
one JSON line: valid frames x NFE per second (padding frames are NOT counted), batches, padding share, the B = 1 loop of the
reference's driver for comparison (same segments, one call per segment)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--segments', type=int, default=96)
    ap.add_argument('--min-s', type=float, default=2.0)
    ap.add_argument('--max-s', type=float, default=30.0)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--max-batch-frames', type=int, default=96 * 128)
    args = ap.parse_args()
    import xiaoicesing_io_b200 as P
    from xiaoicesing_io_b200 import segments as S
    from oracle import weights as OW, denoisers as OD
    dev = torch.device('cuda:0')
    timestep = 512 / 44100
    cfg = OD.WaveNetCfg()
    P.hparams.clear()
    P.hparams.update(hidden_size=cfg.hidden_size, schedule_type='linear', use_shallow_diffusion=False, diff_speedup=50,
                     diff_accelerator='ddim', b2s_precision='fp16')
    model = P.GaussianDiffusion(out_dims=128, num_feats=1, timesteps=1000, k_step=1000, backbone_type='wavenet',
                                backbone_args=dict(num_layers=20, num_channels=256, dilation_cycle_length=4),
                                spec_min=[-12.0], spec_max=[0.0])
    model.denoise_fn.load_state_dict(OW.make_state_dict(cfg, seed=0, sigma_w=0.01), strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(7)
    secs = args.min_s + (args.max_s - args.min_s) * torch.rand(args.segments, generator=g)
    params = [dict(ph_dur=f'{float(s):.4f}', offset=float(i), seed=i) for i, s in enumerate(secs)]
    frames = [S.segment_frames(p, timestep) for p in params]
    conds = {i: torch.randn(frames[i], cfg.hidden_size, generator=g).to(dev) for i in range(args.segments)}
    cond_fn = lambda p, n: (conds[int(p['seed'])], None, None)
    nfe = 20
    batches = S.plan_batches(frames, args.max_batch_frames, 64)
    padded = sum(len(b) * (-(-max(frames[i] for i in b) // 128) * 128) for b in batches)

    def run(**kw):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = S.sample_segments(model, params, cond_fn, timestep, dev, **kw)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    for _ in range(2):
        run(max_batch_frames=args.max_batch_frames)                       # warm-up: captures the graphs of the bucket shapes
    tb = min(run(max_batch_frames=args.max_batch_frames)[0] for _ in range(args.reps))
    for _ in range(2):
        run(max_batch_frames=1, max_batch_size=1)
    t1 = min(run(max_batch_frames=1, max_batch_size=1)[0] for _ in range(args.reps))     # one segment per call, like the reference's driver
    valid = sum(frames)
    print(json.dumps(dict(
        workload=f'{args.segments} segments of {args.min_s:g}-{args.max_s:g} s (uniform), DDIM 20 steps, WaveNet 20x256, fp16, one B200',
        valid_frames=valid, padded_frames=padded, padding_share=1 - valid / padded, batches=len(batches),
        batched=dict(seconds=tb, frame_nfe_per_s=valid * nfe / tb, rtf=tb / (valid * timestep)),
        one_segment_per_call=dict(seconds=t1, frame_nfe_per_s=valid * nfe / t1, rtf=t1 / (valid * timestep)),
        speedup=t1 / tb, note='host wall clock around sample_segments incl. H2D of conditions, noise seeding per segment and D2H of the mels')))


if __name__ == '__main__':
    main()
