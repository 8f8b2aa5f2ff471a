// Force-included into the TSAN harness build.  The two "ranks" of the harness live in ONE process, so the doorbell words of a
// mailbox would be mapped twice (owner's view, sender's view) at two virtual addresses of the same physical page.
// ThreadSanitizer tracks release / acquire pairs per *virtual* address and would not see the sender's store-release and the
// owner's load-acquire as a pair.  Mapping a shm object once per process and handing the same pointer to every mmap() of it
// restores the semantics of the real deployment (two processes, coherent memory, one location).
#pragma once
#include <sys/mman.h>
#include <sys/types.h>

void* harness_mmap(void* addr, size_t len, int prot, int flags, int fd, off_t off);
int harness_munmap(void* addr, size_t len);
#define mmap harness_mmap
#define munmap harness_munmap
