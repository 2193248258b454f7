"""torchrun worker of tests/test_pipeline_gpu.py::test_replicas_bit_identical_across_gpus_nccl: rank 0 materialises the
weights, NCCL broadcast (host/dist.py), every rank runs the same frames through lib.pipeline and the u8 outputs are
all-gathered and compared bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200SD_SYNTHETIC_WEIGHTS", "1")
os.environ["NVENC"] = "1"

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ai_rtc_agent_b200.host import dist as bdist  # noqa: E402
from lib.pipeline import StreamDiffusionPipeline  # noqa: E402


def main():
    rank, world, local = bdist.init()
    dev = torch.device("cuda", local)
    model = os.getenv("REPLICA_MODEL", "tiny-sd15")
    hw = int(os.getenv("REPLICA_HW", "128"))
    bdist.load_and_broadcast(model, dev)
    pipe = StreamDiffusionPipeline(model, t_index_list=[18, 26, 35, 45], width=hw, height=hw)
    g = torch.Generator().manual_seed(7)
    outs = []
    for _ in range(6):
        f = torch.randint(0, 256, (1, hw, hw, 3), dtype=torch.uint8, generator=g).to(dev)
        outs.append(pipe(f))
    mine = torch.stack(outs)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    same = all(torch.equal(gathered[0], t) for t in gathered[1:])
    if rank == 0:
        print(f"replicas identical: {same} (world {world}, {mine.numel()} bytes per rank)")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
