import os, sys, time
os.environ["CATTUS_B200_TRACE_BATCH"]="1"
sys.path.insert(0,'/root/repo')
from cattus_b200 import CudaNetwork
from cattus_b200.export import export_blob
from oracle import games, net
name=sys.argv[1]; streams=int(sys.argv[2])
cfg=net.CONFIGS[name]
n=16384
if cfg.game=='chess': words,bm=games.synth_chess_positions(n,1)
else:
    words,_=games.synth_hex_positions(n,cfg.board_size,1); bm=None
with CudaNetwork(export_blob(net.make_state_dict(cfg,0),cfg.game),cfg.game,batch_size=4096,n_streams=streams) as nw:
    for i in range(3):
        sys.stderr.write(f"--- call {i}\n")
        t0=time.perf_counter(); nw.eval_batch(words,bm); t1=time.perf_counter()
        sys.stderr.write(f"python-side total {1e6*(t1-t0):.1f} us\n")
