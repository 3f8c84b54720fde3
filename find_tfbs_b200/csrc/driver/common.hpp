// common.hpp -- includes, the panic helper and Range shared by every part of the driver.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <map>
#include <unordered_map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/tfbs.h"

namespace {


[[noreturn]] void die(const std::string& msg) {
#ifdef TFBS_DRIVER_TEST_SHIM
    throw std::runtime_error(msg);  // the CPU tests of the host logic catch the "panic"
#else
    fprintf(stderr, "find-tfbs-b200: %s\n", msg.c_str());
    exit(101);  // a Rust panic exits with 101
#endif
}

struct Range {
    uint64_t start, end;  // inclusive (range.rs:4-8)
    bool overlaps(const Range& o) const { return (o.start >= start && o.start <= end) || (o.end >= start && o.end <= end); }  // range.rs:18-21
    bool operator==(const Range& o) const { return start == o.start && end == o.end; }
};

}  // namespace
