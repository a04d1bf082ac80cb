"""Distributed training (data parallel) -- mirror of neural_arch.distributed for the DP subset."""
from .communication import (ProcessGroup, ReduceOp, all_gather, all_reduce, all_reduce_max_host, barrier, broadcast,
                            destroy_process_group, get_backend, get_rank, get_world_size, init_process_group, is_available,
                            is_initialized, recv, reduce_scatter, send)
from .data_parallel import (DataParallel, DistributedDataParallel, DistributedSampler, gather, get_distributed_info,
                            is_distributed_training_available, no_sync, scatter)
from .launcher import DistributedLauncher, LaunchConfig, find_free_port, get_local_ip, launch, launch_distributed_training
