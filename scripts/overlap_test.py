import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench, hpmvs_b200 as hp
scene, desc = bench.cached_scene("plane8", 0)
eng = hp.Engine.from_synth(scene, hp.Options.defaults(), device=0)
seeds, valid = hp.seed_patches(eng.options, eng.cameras, scene.points, scene.meas_offsets, scene.meas_cam)
seeds = np.ascontiguousarray(seeds[valid]); n = len(seeds)
rec = torch.from_numpy(seeds.view(np.uint8).reshape(n, 208)).cuda()
for nstreams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    outs = [torch.empty_like(rec) for _ in range(nstreams)]
    for _ in range(3):
        eng.optimize_device(n, rec.data_ptr(), outs[0].data_ptr(), streams[0].cuda_stream)
    torch.cuda.synchronize(); eng.counters(reset=True)
    K = 12
    t0 = time.perf_counter()
    for k in range(K):
        s = streams[k % nstreams]
        eng.optimize_device(n, rec.data_ptr(), outs[k % nstreams].data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c = eng.counters(reset=True)
    print(f"streams {nstreams}: {1e3*dt/K:.2f} ms/step, {c.patches_ok/dt:.0f} patches/s, ok/step {c.patches_ok/K}")
    ref = outs[0].cpu().numpy()
    for o in outs[1:]:
        assert np.array_equal(o.cpu().numpy(), ref)
