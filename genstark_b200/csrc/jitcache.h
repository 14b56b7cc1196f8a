// On-disk cache shared by the two run-time code generators (hostjit.h: trace loop -> host .so; devjit.cuh: constraint
// kernel -> cubin).  What is loaded from it is executed, so the directory and every file taken from it must belong to
// this user and be writable by nobody else; nothing is ever looked up in a shared /tmp path under a predictable name.
#pragma once
#include <fcntl.h>
#include <spawn.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern char** environ;

namespace gs {

// owner == effective uid, a directory, no group / world write bit
static inline bool jit_dir_trusted(const std::string& dir) {
    struct stat st;
    if (lstat(dir.c_str(), &st) != 0) return false;
    return S_ISDIR(st.st_mode) && st.st_uid == geteuid() && (st.st_mode & 022) == 0;
}
static inline bool jit_file_trusted(const std::string& path) {
    struct stat st;
    if (lstat(path.c_str(), &st) != 0) return false;
    return S_ISREG(st.st_mode) && st.st_uid == geteuid() && (st.st_mode & 022) == 0 && st.st_size > 0;
}

// GS_JIT_CACHE, else $XDG_CACHE_HOME/genstark_b200, else $HOME/.cache/genstark_b200; without any of them (or when the
// directory fails the ownership test) a fresh mkdtemp directory that lives as long as the process.  "" + *why when
// nothing private can be had.
static inline std::string jit_cache_dir(std::string* why) {
    std::string dir, refused;
    if (const char* e = getenv("GS_JIT_CACHE")) dir = e;
    else if (const char* x = getenv("XDG_CACHE_HOME")) dir = std::string(x) + "/genstark_b200";
    else if (const char* h = getenv("HOME")) dir = std::string(h) + "/.cache/genstark_b200";
    if (!dir.empty()) {
        std::string cur;
        for (size_t i = 0; i <= dir.size(); ++i) {
            if (i == dir.size() || dir[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0700); }
            if (i < dir.size()) cur += dir[i];
        }
        if (jit_dir_trusted(dir)) return dir;
        refused = dir + " is not a directory owned by this user that only this user can write to; ";
    }
    static std::string private_tmp;           // one per process
    if (private_tmp.empty()) {
        char tmpl[] = "/tmp/genstark_b200-XXXXXX";
        if (const char* d = mkdtemp(tmpl)) private_tmp = d;
    }
    if (private_tmp.empty() && why) *why = refused + "mkdtemp failed";
    return private_tmp;
}

// run argv[0] (searched in PATH) with stdout + stderr sent to `log`; exit status, or -1
static inline int jit_spawn(const std::vector<std::string>& argv, const std::string& log) {
    std::vector<char*> av;
    for (const std::string& a : argv) av.push_back(const_cast<char*>(a.c_str()));
    av.push_back(nullptr);
    posix_spawn_file_actions_t fa;
    if (posix_spawn_file_actions_init(&fa) != 0) return -1;
    posix_spawn_file_actions_addopen(&fa, 1, log.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0600);
    posix_spawn_file_actions_adddup2(&fa, 1, 2);
    pid_t pid = 0;
    const int rc = posix_spawnp(&pid, av[0], &fa, nullptr, av.data(), environ);
    posix_spawn_file_actions_destroy(&fa);
    if (rc != 0) return -1;
    int status = 0;
    while (waitpid(pid, &status, 0) < 0) { if (errno != EINTR) return -1; }
    return WIFEXITED(status) ? WEXITSTATUS(status) : -1;
}

// what the code was compiled FOR goes into the cache key next to the source: a cache on a shared home must not hand a
// -march=native object to another CPU model
static inline std::string jit_cpu_tag() {
    static std::string tag;
    if (!tag.empty()) return tag;
    std::string model, flags;
    if (FILE* f = fopen("/proc/cpuinfo", "r")) {
        char line[8192];
        while (fgets(line, sizeof line, f)) {
            if (model.empty() && !strncmp(line, "model name", 10)) model = line;
            else if (!strncmp(line, "cpu family", 10) || !strncmp(line, "model\t", 6) || !strncmp(line, "stepping", 8)) model += line;
            else if (flags.empty() && !strncmp(line, "flags", 5)) { flags = line; break; }
        }
        fclose(f);
    }
    tag = model + flags;
    if (tag.empty()) tag = "unknown-cpu";
    return tag;
}

}  // namespace gs
