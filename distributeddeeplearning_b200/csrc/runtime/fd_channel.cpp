// File-descriptor exchange between the ranks of one box (SCM_RIGHTS over abstract unix sockets).
//
// CUDA VMM allocations are shared across processes as POSIX file descriptors
// (cuMemExportToShareableHandle).  A descriptor number means nothing in another process, so it
// has to travel as ancillary data.  Abstract-namespace sockets need no filesystem cleanup and
// no ptrace capability (unlike pidfd_getfd).
#include "host_runtime.h"

#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <cerrno>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace ddl {
namespace {

using Clock = std::chrono::steady_clock;

socklen_t fill_addr(sockaddr_un* addr, const std::string& name) {
  std::memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  if (name.size() + 1 >= sizeof(addr->sun_path)) throw std::runtime_error("fd_channel: socket name too long");
  addr->sun_path[0] = '\0';  // abstract namespace
  std::memcpy(addr->sun_path + 1, name.data(), name.size());
  return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

int make_server(const std::string& name, int backlog) {
  int s = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (s < 0) throw std::runtime_error(std::string("fd_channel: socket(): ") + std::strerror(errno));
  sockaddr_un addr;
  socklen_t len = fill_addr(&addr, name);
  if (::bind(s, reinterpret_cast<sockaddr*>(&addr), len) != 0) {
    int e = errno;
    ::close(s);
    throw std::runtime_error("fd_channel: bind(" + name + "): " + std::strerror(e));
  }
  if (::listen(s, backlog) != 0) {
    int e = errno;
    ::close(s);
    throw std::runtime_error(std::string("fd_channel: listen(): ") + std::strerror(e));
  }
  return s;
}

int connect_retry(const std::string& name, int timeout_ms) {
  auto deadline = Clock::now() + std::chrono::milliseconds(timeout_ms);
  sockaddr_un addr;
  socklen_t len = fill_addr(&addr, name);
  while (true) {
    int s = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (s < 0) throw std::runtime_error(std::string("fd_channel: socket(): ") + std::strerror(errno));
    if (::connect(s, reinterpret_cast<sockaddr*>(&addr), len) == 0) return s;
    ::close(s);
    if (Clock::now() > deadline) throw std::runtime_error("fd_channel: timed out connecting to " + name);
    std::this_thread::sleep_for(std::chrono::milliseconds(2));
  }
}

void send_fd(int sock, int fd, int32_t tag) {
  msghdr msg{};
  iovec iov{&tag, sizeof(tag)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  ssize_t n;
  do { n = ::sendmsg(sock, &msg, MSG_NOSIGNAL); } while (n < 0 && errno == EINTR);
  if (n != static_cast<ssize_t>(sizeof(tag)))
    throw std::runtime_error(std::string("fd_channel: sendmsg(): ") + std::strerror(errno));
}

int recv_fd(int sock, int32_t* tag, int timeout_ms) {
  pollfd p{sock, POLLIN, 0};
  int pr = ::poll(&p, 1, timeout_ms);
  if (pr <= 0) throw std::runtime_error("fd_channel: timed out waiting for a descriptor");
  msghdr msg{};
  iovec iov{tag, sizeof(*tag)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  ssize_t n;
  do { n = ::recvmsg(sock, &msg, MSG_CMSG_CLOEXEC); } while (n < 0 && errno == EINTR);
  if (n != static_cast<ssize_t>(sizeof(*tag)))
    throw std::runtime_error(std::string("fd_channel: recvmsg(): ") + std::strerror(errno));
  for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c != nullptr; c = CMSG_NXTHDR(&msg, c)) {
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      int fd;
      std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
      return fd;
    }
  }
  throw std::runtime_error("fd_channel: message carried no descriptor");
}

// Accept one connection from a process of OUR user (SO_PEERCRED): the socket lives in the abstract namespace, which
// any local process can reach, so connections of other users are dropped instead of being trusted (their bogus
// descriptor or source rank would otherwise poison / abort the arena set-up).  The socket name itself carries a random
// token agreed over the process group (see SymmetricArena), so a foreign process of the same user cannot guess it.
int accept_timeout(int server, int timeout_ms) {
  for (;;) {
    pollfd p{server, POLLIN, 0};
    int pr = ::poll(&p, 1, timeout_ms);
    if (pr <= 0) throw std::runtime_error("fd_channel: timed out in accept");
    int c = ::accept4(server, nullptr, nullptr, SOCK_CLOEXEC);
    if (c < 0) throw std::runtime_error(std::string("fd_channel: accept(): ") + std::strerror(errno));
    ucred cred{};
    socklen_t len = sizeof(cred);
    if (::getsockopt(c, SOL_SOCKET, SO_PEERCRED, &cred, &len) == 0 && cred.uid == ::geteuid()) return c;
    ::close(c);     // not one of ours: ignore and keep waiting for the real peer
  }
}

}  // namespace

std::vector<int> exchange_fds(int rank, int world, int my_fd, const std::string& session,
                              int timeout_ms) {
  std::vector<int> out(world, -1);
  if (world <= 1) return out;
  int server = make_server(session + "-r" + std::to_string(rank), world);
  std::string accept_err;
  std::thread acceptor([&]() {
    try {
      for (int i = 0; i < world - 1; ++i) {
        int c = accept_timeout(server, timeout_ms);
        int32_t src = -1;
        int fd = -1;
        try {
          fd = recv_fd(c, &src, timeout_ms);
        } catch (...) {
          ::close(c);
          throw;
        }
        ::close(c);
        if (src < 0 || src >= world || src == rank) throw std::runtime_error("fd_channel: bad source rank");
        out[src] = fd;
      }
    } catch (const std::exception& e) {
      accept_err = e.what();
    }
  });
  std::string send_err;
  try {
    for (int step = 1; step < world; ++step) {
      int peer = (rank + step) % world;
      int s = connect_retry(session + "-r" + std::to_string(peer), timeout_ms);
      try {
        send_fd(s, my_fd, rank);
      } catch (...) {
        ::close(s);
        throw;
      }
      ::close(s);
    }
  } catch (const std::exception& e) {
    send_err = e.what();
  }
  acceptor.join();
  ::close(server);
  if (!send_err.empty() || !accept_err.empty()) {
    for (int& fd : out)
      if (fd >= 0) { ::close(fd); fd = -1; }
    throw std::runtime_error(send_err.empty() ? accept_err : send_err);
  }
  return out;
}

int broadcast_fd(int rank, int world, int root, int fd, const std::string& session, int timeout_ms) {
  if (world <= 1) return fd;
  if (rank == root) {
    for (int peer = 0; peer < world; ++peer) {
      if (peer == root) continue;
      int s = connect_retry(session + "-b" + std::to_string(peer), timeout_ms);
      try {
        send_fd(s, fd, rank);
      } catch (...) {
        ::close(s);
        throw;
      }
      ::close(s);
    }
    return fd;
  }
  int server = make_server(session + "-b" + std::to_string(rank), 1);
  int got = -1;
  try {
    int c = accept_timeout(server, timeout_ms);
    int32_t src = -1;
    try {
      got = recv_fd(c, &src, timeout_ms);
    } catch (...) {
      ::close(c);
      throw;
    }
    ::close(c);
  } catch (...) {
    ::close(server);
    throw;
  }
  ::close(server);
  return got;
}

}  // namespace ddl
