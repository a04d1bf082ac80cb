"""Distributed training (data parallel) -- mirror of neural_arch.distributed for the DP subset."""
from .communication import (ProcessGroup, ReduceOp, all_gather, all_reduce, all_reduce_max_host, barrier, broadcast,
                            destroy_process_group, get_backend, get_rank, get_world_size, init_process_group, is_available,
                            is_initialized, reduce_scatter)
from .data_parallel import DataParallel, DistributedDataParallel, DistributedSampler
from .launcher import launch
