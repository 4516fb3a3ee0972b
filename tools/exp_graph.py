"""Is one SGCN train step CUDA-graph capturable, and what does it buy on small meshes?  python tools/exp_graph.py [freq ...]"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from semigcn_b200.data import Data
from semigcn_b200.networks import SingleScaleGCN
dev = torch.device('cuda:0')
for freq in [int(a) for a in sys.argv[1:]] or [32, 100]:
    prob = bench.make_problem(freq, dev)
    mesh = prob['mesh']
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv='gcnconv').to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=True)
    data = Data(z1=prob['z1'], x_pos=prob['x_pos'], edge_index=mesh.edge_index)
    dm = prob['dms'][:, 0:1].contiguous()
    def step():
        opt.zero_grad(set_to_none=False)
        out = net(data, dm)
        loss = bench.step_losses(out, prob)
        loss.backward()
        opt.step()
        return loss
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3
    gpu_ms, wall_ms = timeit(step)
    print(f'freq {freq} ({mesh.num_vertices} vertices): eager {gpu_ms:.3f} ms/step (wall {wall_ms:.3f})', flush=True)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = step()
        torch.cuda.synchronize()
        l0 = float(loss)
        gpu_ms, wall_ms = timeit(g.replay)
        print(f'freq {freq}: graph replay {gpu_ms:.3f} ms/step (wall {wall_ms:.3f}), loss {l0:.6f} -> {float(loss):.6f}', flush=True)
    except Exception as e:
        import traceback; traceback.print_exc()
        print('capture failed:', repr(e)[:500])
