import os, sys, time, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time():.1f}", *a, file=sys.stderr, flush=True)
local = int(os.environ["LOCAL_RANK"]); rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
log("init pg")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg ok"); dist.barrier(); log("barrier ok")
from qpalette.decode import DecodeRunner, LLAMA31_8B, uniform_qdict
shape = LLAMA31_8B
r = DecodeRunner(shape, uniform_qdict(shape, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * 32, max_seq=64, rank=rank, world=world, process_group=dist.group.WORLD, num_layers=2)
log("runner built")
r.reset(1); r._step(); torch.cuda.synchronize(); log("eager step ok", r.token.item())
r._step(); torch.cuda.synchronize(); log("eager step 2 ok", r.token.item())
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
if mode == "graph":
    r.capture(); log("capture ok")
    for i in range(4):
        r.step()
    torch.cuda.synchronize(); log("replay ok", r.token.item())
dist.barrier(); torch.cuda.synchronize(); log("done"); sys.stderr.flush(); os._exit(0)
