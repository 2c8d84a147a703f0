"""Seed / hyper-parameter sweeps with the reference's `run_multiple` surface and CSV layout
(run_multiple/final_main_iteration_wb.py:248-262, 1129-1197; run_multiple/final_main_iteration_ca.py:249-255, 1179-1256).

Every member of the sweep (one seed of one (lr, bs, bsr) grid point) is an independent `train_all_epochs` run over the
same GPU-resident embeddings.  The reference runs them one after another in nested Python loops; here
  * the members are dealt round-robin to the ranks of a `torchrun` job (replicated data, no data-path collective --
    SURVEY.md section 8e) and rank 0 gathers the result dictionaries and writes the CSVs;
  * the members of one rank train in LOCK STEP (BASELINE config 5): every member is a `train_all_epochs_gen` generator that
    yields its prepared training epochs; the epochs that share data, batch size, stage and prompts run as ONE
    `dbmm_train_epoch_batched` call (every kernel of the step launched once for all members), the rest one by one.
    Each member keeps its own torch / numpy RNG streams (swapped in around its turns), so it draws the batch orders and
    initial weights of its stand-alone run.  `--sequential_members` restores one-after-another execution.
Per member the result equals the sequential run with that seed, so the mean +- std rows are those of the reference's
protocol -- including its quirk that the `*_std` row is computed after the `*_mean` row was appended (std over seeds +
their mean, ddof = 1).
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from . import cli
from . import engine as E

BLOCKS = (("test", "test"), ("zs_spurious", "zs_spu"), ("train", "tr"), ("val", "val"), ("zs_target", "zs_tg"))


def add_sweep_arguments(parser):
    parser.add_argument("--num_iter", type=int, default=3, help="number of seeds to run")
    parser.add_argument("--random_seeds", type=str, default="42,32,22", help="random seeds, one per iteration")
    parser.add_argument("--lr_multiple", type=float, default=1.0, help="learning_rate_reg = lr_multiple * learning_rate in the grid")
    parser.add_argument("--lr_list", type=str, default=None, help="grid of learning rates (comma separated)")
    parser.add_argument("--bs_list", type=str, default=None, help="grid of batch sizes")
    parser.add_argument("--bsr_list", type=str, default=None, help="grid of stage-2 batch sizes")
    parser.add_argument("--results_root", type=str, default="results_iterative")
    parser.add_argument("--sequential_members", action="store_true", help="run the members one after another (reference order) "
                                                                         "instead of in lock step")
    return parser


def parse_option(argv=None):
    opt = add_sweep_arguments(cli.build_parser()).parse_args(argv)
    opt.random_seeds = [int(t) for t in str(opt.random_seeds).split(",")]
    assert opt.num_iter <= len(opt.random_seeds), "need one seed per iteration"
    for name, typ in (("lr_list", float), ("bs_list", int), ("bsr_list", int)):
        v = getattr(opt, name)
        setattr(opt, name, [typ(t) for t in v.split(",")] if v else None)
    return cli.finalize_options(opt)


def grid_points(opt):
    """(lr, bs, bsr) combinations in the reference's nesting order (lr outermost)."""
    lrs = opt.lr_list or [opt.learning_rate]
    bss = opt.bs_list or [opt.batch_size]
    bsrs = opt.bsr_list or [opt.batch_size_reg]
    return [(lr, bs, bsr) for lr in lrs for bs in bss for bsr in bsrs]


def members(opt):
    """[(grid index, iteration 1..num_iter, seed)] in sequential order."""
    return [(gi, it, opt.random_seeds[it - 1]) for gi in range(len(grid_points(opt))) for it in range(1, opt.num_iter + 1)]


def member_options(opt, point, seed):
    lr, bs, bsr = point
    o = copy.copy(opt)
    o.learning_rate, o.batch_size, o.batch_size_reg, o.random_seed = lr, bs, bsr, seed
    if opt.lr_list is not None:
        o.learning_rate_reg = lr * opt.lr_multiple           # run_multiple/final_main_iteration_ca.py:1183
    return cli.finalize_options(o)                            # re-derives the warm-up constants and seeds the RNGs


def _member_loaders(loaders, point):
    """Loaders over the shared datasets with this grid point's batch sizes (the datasets themselves are not copied)."""
    from .data import EmbeddingLoader
    _, bs, bsr = point
    trainset, train_loader, reg_loader, val_loader, test_loader = loaders

    def clone(ld, batch_size):
        return None if ld is None else EmbeddingLoader(ld.dataset, batch_size=batch_size, shuffle=ld.shuffle)
    bs_val = bsr if reg_loader is not None else bs            # cli.build_loaders: the eval loaders take the stage-2 batch size
    return trainset, clone(train_loader, bs), clone(reg_loader, bsr), clone(val_loader, bs_val), clone(test_loader, bs_val)


def csv_name(opt):
    """run_multiple/final_main_iteration_wb.py:1166-1189."""
    name = f"ds_{opt.dataset}_tl_{opt.tl_method}_bs_{opt.batch_size}_lr_{opt.learning_rate}"
    if "reg" in opt.tl_method:
        name += f"_lrr{opt.learning_rate_reg}_bsr{opt.batch_size_reg}"
        if opt.balance_val:
            name += "_balval"
        if opt.tl_method != "adapter_reg_seq_alter":
            name += "_CP" if opt.use_cls_prompt_in_reg else "_GP"
        if opt.add_adapter:
            name += "_MA" + ("+ni" if opt.init_near_identity else "+rn")
        if opt.continue_from_best and "seq" in opt.tl_method:
            name += "_cont"
    if opt.resample_ce:
        name += "_rs"
    return name + ".csv"


def aggregate(results: dict):
    """results: {iteration: {"train","val","test","zs_target","zs_spurious": metric dict}} -> the reference's DataFrame:
    per block one row per seed, `<blk>_mean`, then `<blk>_std` computed with the mean row already appended."""
    import pandas as pd
    frames = []
    for key, tag in BLOCKS:
        df = pd.concat([pd.DataFrame({k: float(v) for k, v in results[it][key].items()}, index=[it]) for it in sorted(results)])
        df = pd.concat([df, pd.DataFrame(df.mean().to_dict(), index=[f"{tag}_mean"])])
        df = pd.concat([df, pd.DataFrame(df.std().to_dict(), index=[f"{tag}_std"])])
        frames.append(df)
    return pd.concat(frames).round(4)


class _MemberRun:
    """One member in the lock-step driver: its generator and its private RNG streams (torch CPU / CUDA, numpy)."""

    def __init__(self, key, tag, opt, point, seed, loaders):
        self.key, self.tag = key, tag
        self.job, self.result, self.last_run = None, None, None
        self.opt = member_options(opt, point, seed)               # seeds the global RNGs (final_main.py:253) ...
        self._save_rng()                                          # ... which become this member's streams
        self.gen = cli.train_all_epochs_gen(self.opt, loaders=loaders)

    def _save_rng(self):
        self.rng = (torch.get_rng_state(), torch.cuda.get_rng_state() if torch.cuda.is_available() else None, np.random.get_state())

    def _load_rng(self):
        torch.set_rng_state(self.rng[0])
        if self.rng[1] is not None:
            torch.cuda.set_rng_state(self.rng[1])
        np.random.set_state(self.rng[2])

    def advance(self):
        """Run this member up to its next training epoch (-> self.job) or to the end (-> self.result)."""
        self._load_rng()
        E._member_tag = self.tag
        try:
            self.job = next(self.gen) if self.job is None else self.gen.send(None)
        except StopIteration as e:
            self.job, self.result = None, e.value
            self.last_run = getattr(cli.train_all_epochs, "last_run", None)
        finally:
            E._member_tag = None
            self._save_rng()


def run_members_lockstep(runs):
    """Advance all members epoch by epoch; jobs with equal signatures run as one batched call."""
    for r in runs:
        r.advance()
    while True:
        active = [r for r in runs if r.job is not None]
        if not active:
            return
        groups = {}
        for r in active:
            groups.setdefault(r.job.signature(), []).append(r)
        for sig, rs in groups.items():
            if sig is None or len(rs) == 1:
                for r in rs:
                    r.job.run()
            else:
                E.run_jobs_batched([r.job for r in rs])
        for r in active:
            r.advance()


def run_sweep(opt, loaders=None):
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    points = grid_points(opt)
    mine = {}
    todo = [(gi, it, seed) for n, (gi, it, seed) in enumerate(members(opt)) if n % world == rank]
    if getattr(opt, "sequential_members", False) or len(todo) <= 1:
        for gi, it, seed in todo:
            print(f"=============Grid point {points[gi]} iteration {it}/{opt.num_iter} (seed {seed}, rank {rank})=============")
            o = member_options(opt, points[gi], seed)
            (tr, va, te), (zs_t, zs_s) = cli.train_all_epochs(o, loaders=loaders)
            mine[(gi, it)] = dict(train=tr, val=va, test=te, zs_target=zs_t, zs_spurious=zs_s)
    else:
        if loaders is None:                  # one GPU-resident copy of the data for all members of this rank
            loaders = cli.build_loaders(member_options(opt, points[todo[0][0]], todo[0][2]))
        runs = []
        for gi, it, seed in todo:
            print(f"=============Grid point {points[gi]} iteration {it}/{opt.num_iter} (seed {seed}, rank {rank}): lock step=============")
            runs.append(_MemberRun((gi, it), f"m{gi}_{it}", opt, points[gi], seed, _member_loaders(loaders, points[gi])))
        run_members_lockstep(runs)
        for r in runs:
            (tr, va, te), (zs_t, zs_s) = r.result
            mine[r.key] = dict(train=tr, val=va, test=te, zs_target=zs_t, zs_spurious=zs_s)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        mine = {k: v for part in gathered for k, v in part.items()}
    written = []
    if rank == 0:
        os.makedirs(opt.results_root, exist_ok=True)
        for gi, point in enumerate(points):
            o = member_options(opt, point, opt.random_seeds[0])
            df = aggregate({it: mine[(gi, it)] for it in range(1, opt.num_iter + 1)})
            path = os.path.join(opt.results_root, csv_name(o))
            print("Final Results: ", df)
            print("Save to: ", path)
            df.to_csv(path)
            written.append(path)
    return mine, written


def main(argv=None):
    import torch
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    run_sweep(parse_option(argv))
    if dist.is_initialized():
        dist.destroy_process_group()
