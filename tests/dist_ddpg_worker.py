"""Worker of test_two_rank_learner_matches_one_rank_on_the_joint_minibatch: launched by torch.distributed.run, one rank per
GPU, NCCL.  Both ranks hold the same replay contents and parameters, take 3 distributed updates on disjoint rows, and
rank 0 saves what the test compares."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = f"cuda:{torch.cuda.current_device()}"
    import test_gpu_ddpg as T
    from mr_rl_b200.actor import init_actor
    from mr_rl_b200.ddpg import DDPGLearner, ReplayBuffer, init_critic
    rb0, (s, a, r, d, s2) = T.fill_replay(4096, seed=11)          # built on cuda:0 by the helper; rebuild on this rank's GPU
    rb = ReplayBuffer(4096, 0, device=dev)
    for dst, src in ((rb.s, s), (rb.a, a), (rb.r, r), (rb.d, d), (rb.s2, s2)):
        dst.copy_(src)
    rb.count = 4096
    seed = 8
    learner = DDPGLearner(init_actor(seed), init_critic(seed + 1), T.BOUND, actor_target_init=init_actor(seed + 2),
                          critic_target_init=init_critic(seed + 3), device=dev)
    g = torch.Generator().manual_seed(100 + rank)
    used = []
    for u in range(3):
        idx = torch.randperm(2048, generator=g)[:512] + 2048 * rank   # disjoint halves of the ring
        used.append(idx)
        learner.update_distributed(rb, indices=idx.to(dev))
    torch.cuda.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, {"actor": learner.actor.cpu(), "critic": learner.critic.cpu(), "idx": used})
    if rank == 0:
        torch.save({"rank0_actor": gathered[0]["actor"], "rank1_actor": gathered[1]["actor"], "rank0_critic": gathered[0]["critic"],
                    "rank1_critic": gathered[1]["critic"], "idx": [gathered[0]["idx"], gathered[1]["idx"]]}, sys.argv[1])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
