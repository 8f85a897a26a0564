"""One rank of the exchange test as a stand-alone process (RANK / WORLD_SIZE / MASTER_* from the environment, all ranks on GPU 0), so that
every rank can run under compute-sanitizer:
  torchrun --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29544 --no-python \
      compute-sanitizer --tool memcheck python tools/exchange_worker.py /tmp/out"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_exchange as t  # noqa: E402

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
t._worker(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["MASTER_PORT"]), out)
print("rank", os.environ["RANK"], "done")
