// tests/msg_harness.cu — CPU-side check of the fence-free result message protocol (engine.cuh: result_message_consistent): a message
// is accepted only when flag, checksum and every word belong to the same launch, whatever order the stores landed in.
// Test infrastructure: built by tests/test_host_logic.py with nvcc (host code only), never shipped.
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../hdl_graph_slam_b200/csrc/engine.cuh"
namespace b2r { thread_local std::string g_last_error; }
using namespace b2r;

static unsigned long long bits(double d) { unsigned long long u; std::memcpy(&u, &d, 8); return u; }

int main() {
  const int nv = 29;
  unsigned long long s = 99;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) / 9007199254740992.0 - 0.5; };
  long bad = 0, accepted_partial = 0, trials = 0;
  for (int has_extra = 0; has_extra <= 1; has_extra++) {
    std::vector<unsigned long long> mem(nv + 2 + 1, 0);  // words, extra slot, checksum, (+ flag kept separately)
    unsigned long long flag = 0;
    for (unsigned long long seq = 1; seq <= 200; seq++) {
      // the new message, as the device would store it
      std::vector<unsigned long long> msg(nv + 2);
      unsigned long long x = seq;
      for (int i = 0; i < nv; i++) { msg[i] = bits(rnd() * 1e3); x ^= msg_mix(msg[i], i); }
      msg[nv] = has_extra ? (unsigned long long)(seq * 7919) : mem[nv];
      if (has_extra) x ^= msg_mix(msg[nv], nv);
      msg[nv + 1] = x;
      // stores land one by one in a random order; index nv + 2 stands for the flag
      std::vector<int> order(nv + 3);
      for (int i = 0; i < nv + 3; i++) order[i] = i;
      for (int i = nv + 2; i > 0; i--) { const int j = (int)((rnd() + 0.5) * (i + 1)) % (i + 1); std::swap(order[i], order[j]); }
      for (int k = 0; k < nv + 3; k++) {
        // before store k: the message in memory is a mix of launch seq - 1 and launch seq
        const bool ok = result_message_consistent(&flag, seq, mem.data(), nv, has_extra != 0);
        bool complete = flag == seq;
        for (int i = 0; i < nv + 2 && complete; i++) complete = mem[i] == msg[i];
        trials++;
        if (ok && !complete) accepted_partial++;  // would be a torn read (possible only by a 2^-64 accident)
        if (!ok && complete) bad++;
        const int i = order[k];
        if (!has_extra && i == nv) continue;  // no extra word in this message type
        if (i == nv + 2) flag = seq; else mem[i] = msg[i];
      }
      if (!result_message_consistent(&flag, seq, mem.data(), nv, has_extra != 0)) bad++;   // complete message must be accepted
      if (result_message_consistent(&flag, seq + 1, mem.data(), nv, has_extra != 0)) bad++;  // and only for its own sequence number
    }
  }
  {  // two stale words whose old ^ new deltas are equal would cancel in a plain xor checksum; with the position mix they do not
    const int nv2 = 4;
    unsigned long long w[nv2 + 2] = {5, 9, 100, 200, 0, 0};
    unsigned long long flag2 = 7, x = 7;
    for (int i = 0; i < nv2; i++) x ^= msg_mix(w[i], i);
    w[nv2 + 1] = x;
    if (!result_message_consistent(&flag2, 7, w, nv2, false)) bad++;
    unsigned long long stale[nv2 + 2];
    std::memcpy(stale, w, sizeof(w));
    stale[0] ^= 0x33; stale[1] ^= 0x33;  // both words stale by the same delta
    if (result_message_consistent(&flag2, 7, stale, nv2, false)) bad++;
  }
  printf("trials=%ld accepted_partial=%ld rejected_complete=%ld\n", trials, accepted_partial, bad);
  return (bad || accepted_partial) ? 1 : 0;
}
