// MegaDPP transport: asynchronous slot mailbox between pipeline neighbours on one NVSwitch node.
//
// Replaces the reference's host-staged POSIX-shm + semaphore + ibverbs mailbox
// (shm_tensor_new_rdma_pre_alloc.cpp: put_*/get_* :1476-1519, sender/receiver threads :577-801)
// with: payload  = ONE cudaMemcpyAsync into the *receiver-resident* slot (CUDA-IPC mapped, so the
//                  copy engine moves it GPU->GPU over NVLink, any dtype, no host staging);
//       doorbell = a 1-thread kernel on the same stream that does __threadfence_system() and
//                  st.release.sys of the iteration epoch into a flag word that lives in POSIX
//                  shared memory, host-registered in both processes;
//       polling  = the receiver's CPU reads the flag word directly (no CUDA call, no thread,
//                  no semaphore), which is what lets the dynamic scheduler ask "which inputs
//                  are ready right now?" for free.
// Slots are indexed by task (microbatch, chunk) so there is no reuse inside an iteration; the
// epoch value makes flags self-resetting across iterations.
#include "b200_api.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace b200 {

struct Mailbox {
  int n_slots = 0;
  size_t slot_bytes = 0;
  // receive side
  void* recv_buf = nullptr;            // device, owned
  uint32_t* recv_flags_host = nullptr; // shm (host view)
  std::string shm_name;
  int shm_fd = -1;
  size_t flags_bytes = 0;
  // send side
  void* peer_buf = nullptr;            // IPC-opened peer recv_buf
  uint32_t* peer_flags_host = nullptr; // peer's shm mapped here
  uint32_t* peer_flags_dev = nullptr;  // device pointer of the registered mapping
  int peer_fd = -1;
  std::atomic<long> sent{0};
};

static std::map<int, Mailbox*> g_boxes;
static std::mutex g_mu;
static int g_next = 1;

static std::string g_err;
const char* mailbox_last_error() { return g_err.c_str(); }

// the id -> mailbox table is read from the scheduler thread while setup / teardown may run on another one
static Mailbox* find_box(int id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_boxes.find(id);
  return it == g_boxes.end() ? nullptr : it->second;
}

int mailbox_create(int n_slots, size_t slot_bytes, const char* shm_name, void* ipc_handle_out64) {
  auto* mb = new Mailbox();
  mb->n_slots = n_slots;
  mb->slot_bytes = (slot_bytes + 255) / 256 * 256;
  mb->shm_name = shm_name;
  mb->flags_bytes = ((size_t)n_slots * sizeof(uint32_t) + 4095) / 4096 * 4096;
  if (cudaMalloc(&mb->recv_buf, mb->slot_bytes * n_slots) != cudaSuccess) { g_err = "cudaMalloc failed"; delete mb; return -1; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, mb->recv_buf) != cudaSuccess) { g_err = "cudaIpcGetMemHandle failed"; return -1; }
  memcpy(ipc_handle_out64, &h, sizeof(h));
  shm_unlink(shm_name);
  mb->shm_fd = shm_open(shm_name, O_CREAT | O_RDWR, 0600);
  if (mb->shm_fd < 0 || ftruncate(mb->shm_fd, (off_t)mb->flags_bytes) != 0) { g_err = "shm_open/ftruncate failed"; return -1; }
  mb->recv_flags_host = (uint32_t*)mmap(nullptr, mb->flags_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, mb->shm_fd, 0);
  if (mb->recv_flags_host == MAP_FAILED) { g_err = "mmap failed"; return -1; }
  memset(mb->recv_flags_host, 0, mb->flags_bytes);
  std::lock_guard<std::mutex> lk(g_mu);
  int id = g_next++;
  g_boxes[id] = mb;
  return id;
}

int mailbox_connect(int id, const void* peer_ipc_handle64, const char* peer_shm_name) {
  Mailbox* mb = find_box(id);
  if (!mb) { g_err = "unknown mailbox"; return -1; }
  cudaIpcMemHandle_t h;
  memcpy(&h, peer_ipc_handle64, sizeof(h));
  if (cudaIpcOpenMemHandle(&mb->peer_buf, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { g_err = "cudaIpcOpenMemHandle failed"; return -1; }
  mb->peer_fd = shm_open(peer_shm_name, O_RDWR, 0600);
  if (mb->peer_fd < 0) { g_err = std::string("shm_open(peer) failed: ") + peer_shm_name; return -1; }
  mb->peer_flags_host = (uint32_t*)mmap(nullptr, mb->flags_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, mb->peer_fd, 0);
  if (mb->peer_flags_host == MAP_FAILED) { g_err = "mmap(peer) failed"; return -1; }
  if (cudaHostRegister(mb->peer_flags_host, mb->flags_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) { g_err = "cudaHostRegister failed"; return -1; }
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, mb->peer_flags_host, 0) != cudaSuccess) { g_err = "cudaHostGetDevicePointer failed"; return -1; }
  mb->peer_flags_dev = (uint32_t*)d;
  return 0;
}

int mailbox_put(int id, int slot, const void* src, size_t nbytes, uint32_t epoch, cudaStream_t s) {
  Mailbox* mb = find_box(id);
  if (!mb) { g_err = "unknown mailbox"; return -1; }
  if (slot < 0 || slot >= mb->n_slots || nbytes > mb->slot_bytes || !mb->peer_buf) { g_err = "bad slot / size / not connected"; return -1; }
  char* dst = (char*)mb->peer_buf + (size_t)slot * mb->slot_bytes;
  if (cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) { g_err = "peer memcpy failed"; return -1; }
  write_flag(mb->peer_flags_dev + slot, epoch, s);   // doorbell after the payload, same stream
  mb->sent.fetch_add(1);
  return 0;
}

bool mailbox_try_get(int id, int slot, uint32_t epoch) {
  Mailbox* mb = find_box(id);
  if (!mb || slot < 0 || slot >= mb->n_slots) return false;
  // acquire load of the doorbell word the sender's stream wrote (st.release.sys) after the payload
  const uint32_t v = __atomic_load_n(mb->recv_flags_host + slot, __ATOMIC_ACQUIRE);
  return (int32_t)(v - epoch) >= 0 && v != 0;
}

void* mailbox_slot_ptr(int id, int slot) {
  Mailbox* mb = find_box(id);
  if (!mb) return nullptr;
  return (char*)mb->recv_buf + (size_t)slot * mb->slot_bytes;
}

long mailbox_sent(int id) {
  Mailbox* mb = find_box(id);
  return mb ? mb->sent.load() : -1;
}

void mailbox_destroy(int id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_boxes.find(id);
  if (it == g_boxes.end()) return;
  Mailbox* mb = it->second;
  cudaDeviceSynchronize();
  if (mb->peer_flags_host) { cudaHostUnregister(mb->peer_flags_host); munmap(mb->peer_flags_host, mb->flags_bytes); }
  if (mb->peer_fd >= 0) close(mb->peer_fd);
  if (mb->peer_buf) cudaIpcCloseMemHandle(mb->peer_buf);
  if (mb->recv_flags_host) munmap(mb->recv_flags_host, mb->flags_bytes);
  if (mb->shm_fd >= 0) { close(mb->shm_fd); shm_unlink(mb->shm_name.c_str()); }
  if (mb->recv_buf) cudaFree(mb->recv_buf);
  delete mb;
  g_boxes.erase(it);
}

}  // namespace b200
