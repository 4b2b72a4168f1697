// slb_broker — the in-box message broker of the control plane as a native daemon.
//
// The reference depends on an external RabbitMQ server (Erlang) reached through pika (README.md:43-69,
// src/Utils.py:8-32); here the broker is this small C++ process started by server.py / launch.py:
// named FIFO queues, fire-and-forget publish, blocking get with timeout, declare / delete / purge / depth / list.
// One thread per connection (a box has tens of clients, not thousands); payloads are opaque bytes (the clients
// pickle).  Wire protocol (little endian), shared with the Python fallback in transport/broker.py:
//
//   request : u8 op | u32 queue_len | u64 arg_len | queue bytes | arg bytes
//   reply   : u8 status | u64 len | payload            (no reply for PUB)
//   ops     : 1 PUB(arg = body)  2 GET(arg = f64 timeout seconds)  3 DECLARE  4 DELETE  5 PURGE
//             6 DEPTH(-> u64)  7 LIST(-> names joined by '\n')  8 PING  9 SHUTDOWN (loopback peers only)
//             10 AUTH(arg = shared token): when the daemon has a token (environment SLB200_BROKER_TOKEN) this must be
//                the first request of every connection; any other request on an unauthenticated connection closes it.
//                The daemon refuses to bind a non-loopback address without a token.
//
//   slb_broker --host 127.0.0.1 --port 29777      prints "SLB_BROKER_READY <port>" once listening
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>

namespace {

std::string g_token;

bool peer_is_loopback(int fd) {
  sockaddr_in a{};
  socklen_t n = sizeof(a);
  if (::getpeername(fd, reinterpret_cast<sockaddr*>(&a), &n) != 0 || a.sin_family != AF_INET) return false;
  return (ntohl(a.sin_addr.s_addr) >> 24) == 127;
}

bool token_equal(const std::string& a, const std::string& b) {     // constant-time compare
  if (a.size() != b.size()) return false;
  unsigned char d = 0;
  for (size_t i = 0; i < a.size(); ++i) d |= static_cast<unsigned char>(a[i] ^ b[i]);
  return d == 0;
}

struct Store {
  std::mutex m;
  std::condition_variable cv;
  std::unordered_map<std::string, std::deque<std::string>> q;
} g_store;

bool read_exact(int fd, void* buf, size_t n) {
  auto* p = static_cast<uint8_t*>(buf);
  while (n) {
    const ssize_t r = ::recv(fd, p, n, 0);
    if (r <= 0) return false;
    p += r;
    n -= static_cast<size_t>(r);
  }
  return true;
}

bool write_all(int fd, const void* buf, size_t n) {
  auto* p = static_cast<const uint8_t*>(buf);
  while (n) {
    const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
    if (r <= 0) return false;
    p += r;
    n -= static_cast<size_t>(r);
  }
  return true;
}

bool reply(int fd, uint8_t status, const std::string& payload) {
  uint8_t hdr[9];
  hdr[0] = status;
  const uint64_t len = payload.size();
  std::memcpy(hdr + 1, &len, 8);
  return write_all(fd, hdr, 9) && (len == 0 || write_all(fd, payload.data(), len));
}

void serve(int fd) {
  int one = 1;
  ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  bool authed = g_token.empty();
  for (;;) {
    uint8_t hdr[13];
    if (!read_exact(fd, hdr, 13)) break;
    const uint8_t op = hdr[0];
    uint32_t qlen;
    uint64_t alen;
    std::memcpy(&qlen, hdr + 1, 4);
    std::memcpy(&alen, hdr + 5, 8);
    if (qlen > (1u << 16) || alen > (1ull << 33)) break;          // corrupt frame (largest real payload: a ~140 MB state-dict)
    std::string queue, arg;
    try {
      queue.assign(qlen, '\0');
      arg.assign(alen, '\0');
    } catch (const std::bad_alloc&) {
      break;
    }
    if (qlen && !read_exact(fd, &queue[0], qlen)) break;
    if (alen && !read_exact(fd, &arg[0], alen)) break;
    bool ok = true;
    if (op == 10) {                                               // AUTH
      authed = g_token.empty() || token_equal(arg, g_token);
      if (!reply(fd, authed ? 1 : 0, "") || !authed) break;
      continue;
    }
    if (!authed) break;                                           // unauthenticated peer: drop the connection
    switch (op) {
      case 1: {  // PUB
        {
          std::lock_guard<std::mutex> lk(g_store.m);
          g_store.q[queue].emplace_back(std::move(arg));
        }
        g_store.cv.notify_all();
        break;
      }
      case 2: {  // GET
        double timeout = 0.0;
        if (arg.size() == 8) std::memcpy(&timeout, arg.data(), 8);
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout > 0 ? timeout : 0.0);
        std::string body;
        bool have = false;
        {
          std::unique_lock<std::mutex> lk(g_store.m);
          for (;;) {
            auto& dq = g_store.q[queue];
            if (!dq.empty()) {
              body = std::move(dq.front());
              dq.pop_front();
              have = true;
              break;
            }
            if (g_store.cv.wait_until(lk, deadline) == std::cv_status::timeout) {
              auto& dq2 = g_store.q[queue];
              if (!dq2.empty()) {
                body = std::move(dq2.front());
                dq2.pop_front();
                have = true;
              }
              break;
            }
          }
        }
        ok = reply(fd, have ? 1 : 0, body);
        break;
      }
      case 3: {  // DECLARE
        {
          std::lock_guard<std::mutex> lk(g_store.m);
          g_store.q[queue];
        }
        ok = reply(fd, 1, "");
        break;
      }
      case 4: {  // DELETE
        {
          std::lock_guard<std::mutex> lk(g_store.m);
          g_store.q.erase(queue);
        }
        ok = reply(fd, 1, "");
        break;
      }
      case 5: {  // PURGE
        {
          std::lock_guard<std::mutex> lk(g_store.m);
          g_store.q[queue].clear();
        }
        ok = reply(fd, 1, "");
        break;
      }
      case 6: {  // DEPTH
        uint64_t n;
        {
          std::lock_guard<std::mutex> lk(g_store.m);
          n = g_store.q[queue].size();
        }
        ok = reply(fd, 1, std::string(reinterpret_cast<const char*>(&n), 8));
        break;
      }
      case 7: {  // LIST
        std::string names;
        {
          std::lock_guard<std::mutex> lk(g_store.m);
          for (const auto& kv : g_store.q) {
            if (!names.empty()) names.push_back('\n');
            names += kv.first;
          }
        }
        ok = reply(fd, 1, names);
        break;
      }
      case 8:  // PING
        ok = reply(fd, 1, "");
        break;
      case 9:  // SHUTDOWN: only the box itself may stop the broker
        if (!peer_is_loopback(fd)) { ok = false; break; }
        reply(fd, 1, "");
        ::_exit(0);
      default:
        ok = false;
    }
    if (!ok) break;
  }
  ::close(fd);
}

}  // namespace

int main(int argc, char** argv) {
  std::string host = "127.0.0.1";
  int port = 29777;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!std::strcmp(argv[i], "--host")) host = argv[i + 1];
    else if (!std::strcmp(argv[i], "--port")) port = std::atoi(argv[i + 1]);
  }
  if (const char* t = std::getenv("SLB200_BROKER_TOKEN")) g_token = t;
  if (g_token.empty() && host.rfind("127.", 0) != 0) {
    std::fprintf(stderr, "slb_broker: refusing to bind %s without SLB200_BROKER_TOKEN\n", host.c_str());
    return 3;
  }
  ::prctl(PR_SET_PDEATHSIG, SIGTERM);            // never outlive the server process that started us
  ::signal(SIGPIPE, SIG_IGN);
  const int srv = ::socket(AF_INET, SOCK_STREAM, 0);
  if (srv < 0) { std::perror("socket"); return 1; }
  int one = 1;
  ::setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_port = htons(static_cast<uint16_t>(port));
  if (::inet_pton(AF_INET, host.c_str(), &addr.sin_addr) != 1) { std::fprintf(stderr, "bad host %s\n", host.c_str()); return 1; }
  if (::bind(srv, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) { std::perror("bind"); return 2; }
  if (::listen(srv, 128) != 0) { std::perror("listen"); return 2; }
  socklen_t alen = sizeof(addr);
  ::getsockname(srv, reinterpret_cast<sockaddr*>(&addr), &alen);
  std::printf("SLB_BROKER_READY %d\n", ntohs(addr.sin_port));
  std::fflush(stdout);
  for (;;) {
    const int fd = ::accept(srv, nullptr, nullptr);
    if (fd < 0) continue;
    std::thread(serve, fd).detach();
  }
}
