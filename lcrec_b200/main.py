"""Training CLI with the flags of reference ``index/main.py`` (:14-94), including its
``type=bool`` quirk for --bn / --kmeans_init (any non-empty string, e.g. ``--bn False``, is True)."""
import argparse
import logging
import os
import random

import numpy as np
import torch
from torch.utils.data import DataLoader

from .datasets import EmbDataset
from .models.rqvae import RQVAE
from .trainer import Trainer


def build_parser(description="Index"):
    p = argparse.ArgumentParser(description=description)
    p.add_argument("--lr", type=float, default=1e-3)
    p.add_argument("--epochs", type=int, default=5000)
    p.add_argument("--batch_size", type=int, default=2048)
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--eval_step", type=int, default=50)
    p.add_argument("--learner", type=str, default="AdamW")
    p.add_argument("--lr_scheduler_type", type=str, default="constant")
    p.add_argument("--warmup_epochs", type=int, default=50)
    p.add_argument("--data_path", type=str, default="../data/Games/Games.emb-llama-td.npy")
    p.add_argument("--weight_decay", type=float, default=0.0)
    p.add_argument("--dropout_prob", type=float, default=0.0)
    p.add_argument("--bn", type=bool, default=False)
    p.add_argument("--loss_type", type=str, default="mse")
    p.add_argument("--kmeans_init", type=bool, default=True)
    p.add_argument("--kmeans_iters", type=int, default=100)
    p.add_argument("--sk_epsilons", type=float, nargs="+", default=[0.0, 0.0, 0.0])
    p.add_argument("--sk_iters", type=int, default=50)
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--num_emb_list", type=int, nargs="+", default=[256, 256, 256])
    p.add_argument("--e_dim", type=int, default=32)
    p.add_argument("--quant_loss_weight", type=float, default=1.0)
    p.add_argument("--beta", type=float, default=0.25)
    p.add_argument("--layers", type=int, nargs="+", default=[2048, 1024, 512, 256, 128, 64])
    p.add_argument("--save_limit", type=int, default=5)
    p.add_argument("--ckpt_dir", type=str, default="")
    return p


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def main(argv=None):
    seed = 2024
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed); torch.cuda.manual_seed_all(seed)
    args = parse_args(argv)
    print("=================================================")
    print(args)
    print("=================================================")
    logging.basicConfig(level=logging.DEBUG)
    data = EmbDataset(args.data_path)
    model = RQVAE(in_dim=data.dim, num_emb_list=args.num_emb_list, e_dim=args.e_dim, layers=args.layers,
                  dropout_prob=args.dropout_prob, bn=args.bn, loss_type=args.loss_type,
                  quant_loss_weight=args.quant_loss_weight, beta=args.beta, kmeans_init=args.kmeans_init,
                  kmeans_iters=args.kmeans_iters, sk_epsilons=args.sk_epsilons, sk_iters=args.sk_iters)
    print(model)
    loader = DataLoader(data, num_workers=args.num_workers, batch_size=args.batch_size, shuffle=True, pin_memory=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:      # torchrun: one process per GPU, every rank reads the same batches and trains its row block of each
        import torch.distributed as dist
        from .dp_trainer import DataParallelTrainer
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        args.device = f"cuda:{local_rank}"
        dist.init_process_group("nccl")
        trainer = DataParallelTrainer(args, model, len(loader))
    else:
        trainer = Trainer(args, model, len(loader))
    best_loss, best_collision_rate = trainer.fit(loader)
    print("Best Loss", best_loss)
    print("Best Collision Rate", best_collision_rate)


if __name__ == "__main__":
    main()
