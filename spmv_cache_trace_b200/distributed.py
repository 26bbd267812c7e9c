"""Row-partitioned multi-GPU SpMV (one process per GPU, torch.distributed / NCCL).  Filled in below."""


def bench_main(args):
    raise SystemExit("multi-GPU bench not wired yet")
