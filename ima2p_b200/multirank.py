"""One process per GPU: chains shard by rank, loci stay together (they are coupled through the prior,
SURVEY.md fact 1).  The only exchange of an M-mode step is an all-gather of one double per chain --
S = sum_li pdg + probg (swapweight, swapchains.cpp:12-34) -- after which every rank replays the same swap
attempts with the same counter-based random stream and ends with the same beta permutation
(temperatures move, chains do not: swapbetasonly, ima_main_mpi.cpp:1958-1962).  torch.distributed is the
plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class ShardedStepper:
    def __init__(self, engine, device, stream=None):
        self.eng = engine
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.S_local = torch.zeros(engine.nchains, dtype=torch.float64, device=device)
        self.S_global = torch.zeros(engine.nchains_global, dtype=torch.float64, device=device)
        self.stream = stream
        assert engine.nchains * self.world == engine.nchains_global, "chains must shard evenly over ranks"

    def step(self, swaptries):
        self.eng.update_genealogies(self.S_local.data_ptr(), self.stream)
        if self.world > 1:
            dist.all_gather_into_tensor(self.S_global, self.S_local)
        else:
            self.S_global.copy_(self.S_local)
        self.eng.swap_replay(self.S_global.data_ptr(), swaptries, self.stream)

    def run(self, nsteps, swaptries=None):
        st = self.eng.default_swaptries() if swaptries is None else swaptries
        for _ in range(nsteps):
            self.step(st)
