import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sgaligner_b200 import matching
dev = torch.device('cuda:0')
torch.manual_seed(0)
for (cnt, D) in [([[10, 10]], 64), ([[10, 10]], 32), ([[64, 64]], 200), ([[150, 133]], 200)]:
    counts = np.array(cnt)
    N = int(counts.sum())
    emb = torch.randn(N, D).to(dev)
    data = {'graph_per_obj_count': counts}
    a = matching.match_batch(emb, data, k=6, want_sim=True, tensor_cores=True)
    b = matching.match_batch(emb, data, k=6, tensor_cores=False)
    torch.cuda.synchronize()
    sa, sb = a['sim'].view(N, N).cpu(), b['sim'].view(N, N).cpu()
    print(cnt, D, 'max diff', float((sa - sb).abs().max()))
    print(' TC row0:', sa[0, :6].numpy(), '\n FMA row0:', sb[0, :6].numpy())
    print(' TC row5:', sa[5, :6].numpy(), '\n FMA row5:', sb[5, :6].numpy())
    print(' topk tc', a['topk_idx'][0].cpu().numpy(), 'fma', b['topk_idx'][0].cpu().numpy())
