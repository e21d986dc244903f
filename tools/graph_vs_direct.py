import sys, numpy as np
sys.path.insert(0, '/root/repo')
from cattus_b200 import CudaNetwork
from cattus_b200.export import export_blob
from oracle import net, games
cfg = net.CONFIGS["chess10x128"]
w, b = games.synth_chess_positions(4096, 1)
with CudaNetwork(export_blob(net.make_state_dict(cfg, 0), cfg.game), cfg.game, batch_size=4096, n_streams=1) as nw:
    for n in (64, 4096):
        nw.resident_upload(w[:n], b[:n]); nw.time_stage(4, n, 5)
        print(n, "all", float(np.mean(nw.time_stage(4, n, 30))), "split-sum", float(np.mean(nw.time_stage(5, n, 30).sum(axis=1))))
