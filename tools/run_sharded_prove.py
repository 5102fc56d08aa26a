"""Multi-GPU proof under torchrun: every rank proves the same program through the sharded driver over NCCL and checks the
proof against the golden hash (small programs) or against the host verifier (fib19).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_sharded_prove.py"""
import hashlib, importlib, json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
be = pkg.CudaBackend(local)
comm = pkg.Comm.from_torch_distributed(be, dist)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import proof_canon  # noqa: E402
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "proof_hashes.json")))
ok = True
for name in ["with_input", "a-bc", "hello_kakarot", "collatz"]:
    g = GOLD[name]
    src = g["code"].encode() if g["code"] else open(os.path.join(ROOT, "tests", "golden", "programs", name + ".bf"), "rb").read()
    proof = pkg.prove_brainfuck_sharded(be, comm, src, bytes.fromhex(g["stdin_hex"]), g["log_max_rows"])
    proof.verify()
    same = hashlib.sha256(proof_canon.canonical(proof.json().encode())).hexdigest() == g["sha256"]
    ok &= same
    if rank == 0:
        print(json.dumps({"program": name, "world": world, "matches_golden": same}))
code = open(os.path.join(ROOT, "tests", "golden", "programs", "fib19.bf"), "rb").read()
n_fib = int(sys.argv[sys.argv.index("--fib19-proofs") + 1]) if "--fib19-proofs" in sys.argv else 4
for it in range(n_fib):
    dist.barrier()
    t = time.time()
    # the first proof sizes the receive windows of the direct exchange (NCCL all-to-all meanwhile), the later ones use them;
    # odd iterations run the VM and the tables in front of the device work (SBF_NO_OVERLAP, bench.py's `value` timing)
    proof = pkg.prove_brainfuck_sharded(be, comm, code, b"", 24, overlap_host=(it % 2 == 0))
    dt = time.time() - t
    proof.verify()
    h = hashlib.sha256(proof.json().encode()).hexdigest()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    if rank == 0:
        print(json.dumps({"program": "fib19", "world": world, "wall_s": dt, "sha256": h, "ranks_agree": len(set(hs)) == 1, **proof.report()}))
comm.close()
be.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
