// ThreadSanitizer harness for the MegaDPP mailbox (megatronapp_b200/csrc/mailbox.cpp) against the host-only CUDA shim.
// Two "ranks" (threads) own one mailbox each and stream N slots to each other per iteration while a third thread keeps
// creating and destroying unrelated mailboxes (table mutation under load); receivers poll mailbox_try_get exactly like the
// dynamic pipeline scheduler does and verify every payload.  Build: see tests/test_native_sanitizers.py.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "b200_api.h"

#undef mmap
#undef munmap
#include <sys/stat.h>

#include <map>
#include <mutex>

// one mapping per shm object per process (see fake_cuda/shm_alias.h)
namespace {
struct Alias { void* addr; size_t len; int refs; };
std::map<ino_t, Alias> g_alias;
std::mutex g_alias_mu;
}  // namespace

void* harness_mmap(void* addr, size_t len, int prot, int flags, int fd, off_t off) {
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0) return mmap(addr, len, prot, flags, fd, off);
  std::lock_guard<std::mutex> g(g_alias_mu);
  auto it = g_alias.find(st.st_ino);
  if (it != g_alias.end()) { it->second.refs++; return it->second.addr; }
  void* p = mmap(addr, len, prot, flags, fd, off);
  if (p != MAP_FAILED) g_alias[st.st_ino] = {p, len, 1};
  return p;
}

int harness_munmap(void* addr, size_t len) {
  std::lock_guard<std::mutex> g(g_alias_mu);
  for (auto it = g_alias.begin(); it != g_alias.end(); ++it)
    if (it->second.addr == addr) {
      if (--it->second.refs > 0) return 0;
      g_alias.erase(it);
      break;
    }
  return munmap(addr, len);
}

namespace b200 {
// the doorbell kernel of comm.cu: release-store of the epoch after the payload, on the same "stream"
void write_flag(uint32_t* flag, uint32_t value, cudaStream_t) { __atomic_store_n(flag, value, __ATOMIC_RELEASE); }
}  // namespace b200

using namespace b200;

constexpr int kSlots = 16, kWords = 256, kIters = 200;

struct Rank {
  int id = -1;
  char handle[64];
  std::string shm;
};

// the optimizer step / gradient all-reduce that separates two pipeline iterations: no rank starts filling iteration it+1's
// slots while a peer may still be reading iteration it's payloads
static std::atomic<int> g_arrived{0};
static void iteration_barrier(int it) {
  g_arrived.fetch_add(1, std::memory_order_acq_rel);
  while (g_arrived.load(std::memory_order_acquire) < 2 * it) std::this_thread::yield();
}

static void worker(Rank* me, int who, std::atomic<int>* errors) {
  std::vector<uint32_t> src(kWords);
  for (int it = 1; it <= kIters; ++it) {
    for (int s = 0; s < kSlots; ++s) {
      for (int w = 0; w < kWords; ++w) src[w] = (uint32_t)(who * 1000003 + it * 1009 + s * 17 + w);
      if (mailbox_put(me->id, s, src.data(), kWords * 4, (uint32_t)it, nullptr) != 0) errors->fetch_add(1);
    }
    for (int s = 0; s < kSlots; ++s) {
      while (!mailbox_try_get(me->id, s, (uint32_t)it)) std::this_thread::yield();
      const uint32_t* p = (const uint32_t*)mailbox_slot_ptr(me->id, s);
      const int other = 1 - who;
      const uint32_t base = p[0] - (uint32_t)(other * 1000003 + s * 17);
      if (base % 1009 != 0 || base / 1009 != (uint32_t)it) errors->fetch_add(1);
      for (int w = 1; w < kWords; ++w)
        if (p[w] != p[0] + (uint32_t)w) { errors->fetch_add(1); break; }
    }
    iteration_barrier(it);
  }
}

int main() {
  Rank r[2];
  for (int i = 0; i < 2; ++i) {
    r[i].shm = "/b200_tsan_" + std::to_string((long)getpid()) + "_" + std::to_string(i);
    r[i].id = mailbox_create(kSlots, kWords * 4, r[i].shm.c_str(), r[i].handle);
    if (r[i].id < 0) { printf("create failed: %s\n", mailbox_last_error()); return 2; }
  }
  for (int i = 0; i < 2; ++i)
    if (mailbox_connect(r[i].id, r[1 - i].handle, r[1 - i].shm.c_str()) != 0) { printf("connect failed: %s\n", mailbox_last_error()); return 2; }
  std::atomic<int> errors{0};
  std::atomic<bool> stop{false};
  std::thread churn([&] {
    char h[64];
    int n = 0;
    while (!stop.load()) {
      std::string name = "/b200_tsan_churn_" + std::to_string((long)getpid()) + "_" + std::to_string(n++ % 4);
      int id = mailbox_create(2, 64, name.c_str(), h);
      if (id > 0) mailbox_destroy(id);
    }
  });
  std::thread a(worker, &r[0], 0, &errors), b(worker, &r[1], 1, &errors);
  a.join();
  b.join();
  stop.store(true);
  churn.join();
  const long sent = mailbox_sent(r[0].id) + mailbox_sent(r[1].id);
  for (int i = 0; i < 2; ++i) mailbox_destroy(r[i].id);
  printf("mailbox_tsan: sent=%ld errors=%d\n", sent, errors.load());
  return (errors.load() == 0 && sent == 2L * kSlots * kIters) ? 0 : 1;
}
