"""Import-compatible stand-in for the reference's ``shm_tensor_new_rdma`` C++ extension (see megatronapp_b200/dpp/compat_transport.py)."""
from megatronapp_b200.dpp.compat_transport import *  # noqa: F401,F403
from megatronapp_b200.dpp.compat_transport import (check_send_finish, clean_rdma, get_backward_tensor, get_forward_tensor,  # noqa: F401
                                                   init_backward_rdma, init_forward_rdma, init_shared_memory, join_threads,
                                                   put_backward_tensor, put_forward_tensor, thread_pool)
