// Native reader / writer for the tensor files of torbi's file API
// (/root/reference/torbi/core.py:211-368: torch.load of one (frames, states)
// posteriorgram per input file, torch.save of one int32 index array per output
// file; /root/reference/torbi/data/dataset.py:17-22, collate.py:9-33).
//
// torch.save writes a zip container whose entries are STORED (no compression):
// `<name>/data.pkl` (a protocol-2 pickle describing the tensor), `<name>/data/<key>`
// (the raw storage bytes) and a few bookkeeping records.  Here the central directory
// and the small pickle are parsed in C++ and the payload is read with pread()
// straight into the caller's (pinned) batch buffer by a pool of host threads -- no
// Python, no GIL, no intermediate copy.  Anything that is not the plain layout
// (compressed or zip64 archives, several storages, non-contiguous views, unknown
// pickle opcodes, legacy non-zip files) is reported per file through `status`, and
// the Python side falls back to torch.load for that file.
//
// The writer emits the same container for an int32 tensor so that torch.load
// (including weights_only=True) reads it back.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/torbi_b200.h"

namespace {

// ------------------------------------------------------------------ helpers
bool read_at(int fd, void *buffer, size_t bytes, uint64_t offset) {
    char *p = static_cast<char *>(buffer);
    while (bytes > 0) {
        const ssize_t got = pread(fd, p, std::min<size_t>(bytes, size_t(1) << 30), (off_t)offset);
        if (got <= 0) return false;
        p += got;
        offset += (uint64_t)got;
        bytes -= (size_t)got;
    }
    return true;
}

uint16_t u16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t u32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

struct FileDescriptor {                  // closes on every path, exceptions included
    int fd;
    explicit FileDescriptor(int fd_) : fd(fd_) {}
    ~FileDescriptor() { if (fd >= 0) close(fd); }
    FileDescriptor(const FileDescriptor &) = delete;
    FileDescriptor &operator=(const FileDescriptor &) = delete;
};

struct Entry {
    std::string name;
    uint64_t local_offset = 0, size = 0;
    uint16_t method = 0;
};

// central directory of a (non-zip64) archive
int list_entries(int fd, uint64_t file_size, std::vector<Entry> *entries) {
    if (file_size < 22) return TORBI_B200_PT_NOT_A_ZIP;
    const size_t tail_bytes = (size_t)std::min<uint64_t>(file_size, 65536 + 22);
    std::vector<unsigned char> tail(tail_bytes);
    if (!read_at(fd, tail.data(), tail_bytes, file_size - tail_bytes)) return TORBI_B200_PT_IO_ERROR;
    int64_t eocd = -1;
    for (int64_t i = (int64_t)tail_bytes - 22; i >= 0; --i)
        if (u32(&tail[i]) == 0x06054b50u) { eocd = i; break; }
    if (eocd < 0) return TORBI_B200_PT_NOT_A_ZIP;
    const unsigned char *e = &tail[eocd];
    const uint32_t count = u16(e + 10), cd_size = u32(e + 12), cd_offset = u32(e + 16);
    if (count == 0xffffu || cd_size == 0xffffffffu || cd_offset == 0xffffffffu) return TORBI_B200_PT_UNSUPPORTED;  // zip64
    if ((uint64_t)cd_offset + cd_size > file_size) return TORBI_B200_PT_NOT_A_ZIP;
    std::vector<unsigned char> cd(cd_size);
    if (!read_at(fd, cd.data(), cd_size, cd_offset)) return TORBI_B200_PT_IO_ERROR;
    size_t at = 0;
    for (uint32_t k = 0; k < count; ++k) {
        if (at + 46 > cd.size() || u32(&cd[at]) != 0x02014b50u) return TORBI_B200_PT_NOT_A_ZIP;
        Entry entry;
        entry.method = u16(&cd[at + 10]);
        const uint32_t compressed = u32(&cd[at + 20]), size = u32(&cd[at + 24]);
        const uint16_t name_len = u16(&cd[at + 28]), extra_len = u16(&cd[at + 30]), comment_len = u16(&cd[at + 32]);
        const uint32_t local = u32(&cd[at + 42]);
        if (size == 0xffffffffu || compressed == 0xffffffffu || local == 0xffffffffu) return TORBI_B200_PT_UNSUPPORTED;
        if (at + 46 + name_len > cd.size()) return TORBI_B200_PT_NOT_A_ZIP;
        entry.name.assign(reinterpret_cast<const char *>(&cd[at + 46]), name_len);
        entry.size = size;
        entry.local_offset = local;
        if (entry.method == 0 && compressed != size) return TORBI_B200_PT_NOT_A_ZIP;
        entries->push_back(std::move(entry));
        at += 46u + name_len + extra_len + comment_len;
    }
    return TORBI_B200_PT_OK;
}

// where the bytes of a stored entry start
int entry_data_offset(int fd, const Entry &entry, uint64_t *offset) {
    unsigned char header[30];
    if (!read_at(fd, header, sizeof(header), entry.local_offset)) return TORBI_B200_PT_IO_ERROR;
    if (u32(header) != 0x04034b50u) return TORBI_B200_PT_NOT_A_ZIP;
    *offset = entry.local_offset + 30u + u16(header + 26) + u16(header + 28);
    return TORBI_B200_PT_OK;
}

// --------------------------------------------------- the tensor pickle (protocol 2-4 subset)
struct Value {
    enum Kind { kInt, kString, kGlobal, kTuple, kMark, kPersistent, kTensor, kOther } kind = kOther;
    int64_t number = 0;
    std::string text;
    std::vector<Value> items;
};

struct TensorRecord {
    std::string storage_class, key;
    int64_t storage_numel = 0, offset = 0;
    std::vector<int64_t> shape, stride;
    bool found = false;
};

int parse_tensor_pickle(const std::vector<unsigned char> &pkl, TensorRecord *out) {
    std::vector<Value> stack;
    std::vector<Value> memo;
    size_t at = 0;
    auto need = [&](size_t n) { return at + n <= pkl.size(); };
    auto pop_to_mark = [&](std::vector<Value> *items) {
        while (!stack.empty() && stack.back().kind != Value::kMark) {
            items->push_back(std::move(stack.back()));
            stack.pop_back();
        }
        if (stack.empty()) return false;
        stack.pop_back();
        std::reverse(items->begin(), items->end());
        return true;
    };
    auto read_line = [&](std::string *line) {
        const size_t start = at;
        while (at < pkl.size() && pkl[at] != '\n') ++at;
        if (at >= pkl.size()) return false;
        line->assign(reinterpret_cast<const char *>(&pkl[start]), at - start);
        ++at;
        return true;
    };
    auto make_tuple = [&](size_t n) {
        if (stack.size() < n) return false;
        Value tuple;
        tuple.kind = Value::kTuple;
        tuple.items.assign(std::make_move_iterator(stack.end() - n), std::make_move_iterator(stack.end()));
        stack.resize(stack.size() - n);
        stack.push_back(std::move(tuple));
        return true;
    };
    auto memoize = [&](size_t index) {
        if (stack.empty() || index > 4096) return false;           // (a tensor record memoizes a dozen values)
        if (memo.size() <= index) memo.resize(index + 1);
        memo[index] = stack.back();
        return true;
    };
    while (at < pkl.size()) {
        const unsigned char op = pkl[at++];
        Value v;
        switch (op) {
            case 0x80: if (!need(1)) return TORBI_B200_PT_UNSUPPORTED; ++at; break;             // PROTO
            case 0x95: if (!need(8)) return TORBI_B200_PT_UNSUPPORTED; at += 8; break;          // FRAME
            case 'c': {                                                                            // GLOBAL
                std::string module, name;
                if (!read_line(&module) || !read_line(&name)) return TORBI_B200_PT_UNSUPPORTED;
                v.kind = Value::kGlobal;
                v.text = module + " " + name;
                stack.push_back(std::move(v));
                break;
            }
            case 0x93: {                                                                           // STACK_GLOBAL
                if (stack.size() < 2 || stack[stack.size() - 1].kind != Value::kString ||
                    stack[stack.size() - 2].kind != Value::kString) return TORBI_B200_PT_UNSUPPORTED;
                v.kind = Value::kGlobal;
                v.text = stack[stack.size() - 2].text + " " + stack[stack.size() - 1].text;
                stack.resize(stack.size() - 2);
                stack.push_back(std::move(v));
                break;
            }
            case 'q': if (!need(1) || !memoize(pkl[at])) return TORBI_B200_PT_UNSUPPORTED; ++at; break;          // BINPUT
            case 'r': if (!need(4) || !memoize(u32(&pkl[at]))) return TORBI_B200_PT_UNSUPPORTED; at += 4; break;   // LONG_BINPUT
            case 0x94: if (!memoize(memo.size())) return TORBI_B200_PT_UNSUPPORTED; break;                          // MEMOIZE
            case 'h': if (!need(1) || pkl[at] >= memo.size()) return TORBI_B200_PT_UNSUPPORTED; stack.push_back(memo[pkl[at]]); ++at; break;  // BINGET
            case 'j': if (!need(4) || u32(&pkl[at]) >= memo.size()) return TORBI_B200_PT_UNSUPPORTED; stack.push_back(memo[u32(&pkl[at])]); at += 4; break;
            case '(': v.kind = Value::kMark; stack.push_back(std::move(v)); break;                 // MARK
            case 'X': {                                                                            // BINUNICODE
                if (!need(4)) return TORBI_B200_PT_UNSUPPORTED;
                const uint32_t n = u32(&pkl[at]);
                at += 4;
                if (!need(n)) return TORBI_B200_PT_UNSUPPORTED;
                v.kind = Value::kString;
                v.text.assign(reinterpret_cast<const char *>(&pkl[at]), n);
                at += n;
                stack.push_back(std::move(v));
                break;
            }
            case 0x8c: {                                                                           // SHORT_BINUNICODE
                if (!need(1)) return TORBI_B200_PT_UNSUPPORTED;
                const uint32_t n = pkl[at++];
                if (!need(n)) return TORBI_B200_PT_UNSUPPORTED;
                v.kind = Value::kString;
                v.text.assign(reinterpret_cast<const char *>(&pkl[at]), n);
                at += n;
                stack.push_back(std::move(v));
                break;
            }
            case 'K': if (!need(1)) return TORBI_B200_PT_UNSUPPORTED; v.kind = Value::kInt; v.number = pkl[at]; at += 1; stack.push_back(v); break;
            case 'M': if (!need(2)) return TORBI_B200_PT_UNSUPPORTED; v.kind = Value::kInt; v.number = u16(&pkl[at]); at += 2; stack.push_back(v); break;
            case 'J': if (!need(4)) return TORBI_B200_PT_UNSUPPORTED; v.kind = Value::kInt; v.number = (int32_t)u32(&pkl[at]); at += 4; stack.push_back(v); break;
            case 0x8a: {                                                                           // LONG1
                if (!need(1)) return TORBI_B200_PT_UNSUPPORTED;
                const uint32_t n = pkl[at++];
                if (n > 8 || !need(n)) return TORBI_B200_PT_UNSUPPORTED;
                uint64_t bits = 0;
                for (uint32_t k = 0; k < n; ++k) bits |= (uint64_t)pkl[at + k] << (8 * k);
                if (n > 0 && n < 8 && (pkl[at + n - 1] & 0x80)) bits |= ~uint64_t(0) << (8 * n);
                at += n;
                v.kind = Value::kInt;
                v.number = (int64_t)bits;
                stack.push_back(v);
                break;
            }
            case 't': {                                                                            // TUPLE
                v.kind = Value::kTuple;
                if (!pop_to_mark(&v.items)) return TORBI_B200_PT_UNSUPPORTED;
                stack.push_back(std::move(v));
                break;
            }
            case ')': v.kind = Value::kTuple; stack.push_back(std::move(v)); break;                // EMPTY_TUPLE
            case 0x85: if (!make_tuple(1)) return TORBI_B200_PT_UNSUPPORTED; break;
            case 0x86: if (!make_tuple(2)) return TORBI_B200_PT_UNSUPPORTED; break;
            case 0x87: if (!make_tuple(3)) return TORBI_B200_PT_UNSUPPORTED; break;
            case 0x88: case 0x89: case 'N': case '}': case ']': stack.push_back(v); break;         // True / False / None / {} / []
            case 'Q': {                                                                            // BINPERSID
                if (stack.empty() || stack.back().kind != Value::kTuple) return TORBI_B200_PT_UNSUPPORTED;
                stack.back().kind = Value::kPersistent;
                break;
            }
            case 'R': {                                                                            // REDUCE
                if (stack.size() < 2) return TORBI_B200_PT_UNSUPPORTED;
                Value args = std::move(stack.back());
                stack.pop_back();
                Value callable = std::move(stack.back());
                stack.pop_back();
                if (callable.kind != Value::kGlobal || args.kind != Value::kTuple) return TORBI_B200_PT_UNSUPPORTED;
                Value result;
                if (callable.text == "torch._utils _rebuild_tensor_v2") {
                    if (out->found) return TORBI_B200_PT_UNSUPPORTED;                              // more than one tensor
                    if (args.items.size() < 4) return TORBI_B200_PT_UNSUPPORTED;
                    const Value &storage = args.items[0];
                    if (storage.kind != Value::kPersistent || storage.items.size() < 5 ||
                        storage.items[0].kind != Value::kString || storage.items[0].text != "storage" ||
                        storage.items[1].kind != Value::kGlobal || storage.items[2].kind != Value::kString ||
                        storage.items[4].kind != Value::kInt || args.items[1].kind != Value::kInt ||
                        args.items[2].kind != Value::kTuple || args.items[3].kind != Value::kTuple)
                        return TORBI_B200_PT_UNSUPPORTED;
                    out->storage_class = storage.items[1].text;
                    out->key = storage.items[2].text;
                    out->storage_numel = storage.items[4].number;
                    out->offset = args.items[1].number;
                    for (const Value &d : args.items[2].items) {
                        if (d.kind != Value::kInt) return TORBI_B200_PT_UNSUPPORTED;
                        out->shape.push_back(d.number);
                    }
                    for (const Value &d : args.items[3].items) {
                        if (d.kind != Value::kInt) return TORBI_B200_PT_UNSUPPORTED;
                        out->stride.push_back(d.number);
                    }
                    out->found = true;
                    result.kind = Value::kTensor;
                } else if (callable.text != "collections OrderedDict") {
                    return TORBI_B200_PT_UNSUPPORTED;
                }
                stack.push_back(std::move(result));
                break;
            }
            case '.':                                                                              // STOP
                if (stack.size() != 1 || stack.back().kind != Value::kTensor) return TORBI_B200_PT_UNSUPPORTED;
                return TORBI_B200_PT_OK;
            default:
                return TORBI_B200_PT_UNSUPPORTED;
        }
    }
    return TORBI_B200_PT_UNSUPPORTED;
}

int dtype_of(const std::string &storage_class, int *itemsize) {
    struct { const char *name; int dtype, bytes; } table[] = {
        {"torch FloatStorage", TORBI_B200_PT_F32, 4}, {"torch HalfStorage", TORBI_B200_PT_F16, 2},
        {"torch BFloat16Storage", TORBI_B200_PT_BF16, 2}, {"torch DoubleStorage", TORBI_B200_PT_F64, 8},
        {"torch IntStorage", TORBI_B200_PT_I32, 4}, {"torch LongStorage", TORBI_B200_PT_I64, 8},
    };
    for (const auto &row : table)
        if (storage_class == row.name) { *itemsize = row.bytes; return row.dtype; }
    return -1;
}

bool ends_with(const std::string &text, const std::string &suffix) {
    return text.size() >= suffix.size() && text.compare(text.size() - suffix.size(), suffix.size(), suffix) == 0;
}

int inspect_one_unchecked(const char *path, torbi_b200_pt_info *info);

// no C++ exception may cross the C ABI: anything unexpected (allocation failure on a hostile size
// field, ...) makes the file "unsupported" and the caller falls back to torch.load
int inspect_one(const char *path, torbi_b200_pt_info *info) {
    try {
        return inspect_one_unchecked(path, info);
    } catch (...) {
        std::memset(info, 0, sizeof(*info));
        info->status = TORBI_B200_PT_UNSUPPORTED;
        return TORBI_B200_PT_UNSUPPORTED;
    }
}

int inspect_one_unchecked(const char *path, torbi_b200_pt_info *info) {
    std::memset(info, 0, sizeof(*info));
    const FileDescriptor file(open(path, O_RDONLY | O_CLOEXEC));
    const int fd = file.fd;
    if (fd < 0) {
        info->status = TORBI_B200_PT_IO_ERROR;
        return TORBI_B200_PT_IO_ERROR;
    }
    struct stat st;
    int rc = fstat(fd, &st) == 0 ? TORBI_B200_PT_OK : TORBI_B200_PT_IO_ERROR;
    std::vector<Entry> entries;
    if (rc == TORBI_B200_PT_OK) rc = list_entries(fd, (uint64_t)st.st_size, &entries);
    const Entry *pkl = nullptr, *byteorder = nullptr;
    int storages = 0;
    if (rc == TORBI_B200_PT_OK) {
        for (const Entry &entry : entries) {
            if (ends_with(entry.name, "/data.pkl")) pkl = &entry;
            else if (ends_with(entry.name, "/byteorder")) byteorder = &entry;
            else if (entry.name.find("/data/") != std::string::npos) ++storages;
        }
        if (!pkl || storages != 1 || pkl->method != 0 || pkl->size > (1u << 20)) rc = TORBI_B200_PT_UNSUPPORTED;
    }
    if (rc == TORBI_B200_PT_OK && byteorder) {
        uint64_t at = 0;
        char text[8] = {0};
        rc = entry_data_offset(fd, *byteorder, &at);
        if (rc == TORBI_B200_PT_OK && (byteorder->size != 6 || !read_at(fd, text, 6, at) || std::memcmp(text, "little", 6)))
            rc = TORBI_B200_PT_UNSUPPORTED;
    }
    TensorRecord tensor;
    if (rc == TORBI_B200_PT_OK) {
        uint64_t at = 0;
        rc = entry_data_offset(fd, *pkl, &at);
        std::vector<unsigned char> bytes(pkl->size);
        if (rc == TORBI_B200_PT_OK && !read_at(fd, bytes.data(), bytes.size(), at)) rc = TORBI_B200_PT_IO_ERROR;
        if (rc == TORBI_B200_PT_OK) rc = parse_tensor_pickle(bytes, &tensor);
    }
    if (rc == TORBI_B200_PT_OK) {
        int itemsize = 0;
        const int dtype = dtype_of(tensor.storage_class, &itemsize);
        const Entry *storage = nullptr;
        for (const Entry &entry : entries)
            if (ends_with(entry.name, "/data/" + tensor.key)) storage = &entry;
        if (dtype < 0 || !storage || storage->method != 0 || tensor.shape.size() > 4 ||
            tensor.shape.size() != tensor.stride.size() || tensor.offset < 0) {
            rc = TORBI_B200_PT_UNSUPPORTED;
        } else {
            // contiguous (row-major) views only
            int64_t numel = 1, expect = 1;
            for (int d = (int)tensor.shape.size() - 1; d >= 0; --d) {
                // (bounded so that the products below cannot overflow)
                if (tensor.shape[d] < 0 || tensor.shape[d] > (int64_t(1) << 40) || numel > (int64_t(1) << 20) * (int64_t(1) << 20) / std::max<int64_t>(tensor.shape[d], 1)) {
                    rc = TORBI_B200_PT_UNSUPPORTED;
                    break;
                }
                if (tensor.shape[d] > 1 && tensor.stride[d] != expect) rc = TORBI_B200_PT_UNSUPPORTED;
                expect *= std::max<int64_t>(tensor.shape[d], 1);
                numel *= tensor.shape[d];
            }
            uint64_t at = 0;
            if (rc == TORBI_B200_PT_OK) rc = entry_data_offset(fd, *storage, &at);
            if (rc == TORBI_B200_PT_OK &&
                (tensor.offset > (int64_t(1) << 40) || (uint64_t)(tensor.offset + numel) * itemsize > storage->size ||
                 at + storage->size > (uint64_t)st.st_size))
                rc = TORBI_B200_PT_UNSUPPORTED;
            if (rc == TORBI_B200_PT_OK) {
                info->data_offset = (int64_t)at + tensor.offset * itemsize;
                info->nbytes = numel * itemsize;
                info->ndim = (int32_t)tensor.shape.size();
                info->dtype = dtype;
                for (size_t d = 0; d < tensor.shape.size(); ++d) info->shape[d] = tensor.shape[d];
            }
        }
    }
    info->status = rc;
    return rc;
}

template <typename Fn>
void parallel_for(int64_t n, int threads, Fn fn) {
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), n));
    if (threads == 1) {
        for (int64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; ++k)
        pool.emplace_back([&] {
            for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto &worker : pool) worker.join();
}

// ------------------------------------------------------------------ writer
uint32_t crc32_of(const unsigned char *data, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static std::atomic<bool> ready{false};
    if (!ready.load(std::memory_order_acquire)) {
        uint32_t local[256];
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            local[i] = c;
        }
        std::memcpy(table, local, sizeof(table));       // idempotent: racing threads write the same bytes
        ready.store(true, std::memory_order_release);
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ data[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

void put16(std::string *out, uint16_t v) { out->push_back((char)(v & 0xff)); out->push_back((char)(v >> 8)); }
void put32(std::string *out, uint32_t v) { put16(out, (uint16_t)(v & 0xffff)); put16(out, (uint16_t)(v >> 16)); }

void pickle_int(std::string *out, int64_t v) {
    if (v >= 0 && v < 256) { out->push_back('K'); out->push_back((char)v); }
    else if (v >= 0 && v < 65536) { out->push_back('M'); put16(out, (uint16_t)v); }
    else if (v >= INT32_MIN && v <= INT32_MAX) { out->push_back('J'); put32(out, (uint32_t)(int32_t)v); }
    else { out->push_back((char)0x8a); out->push_back(8); for (int k = 0; k < 8; ++k) out->push_back((char)((uint64_t)v >> (8 * k))); }
}

void pickle_tuple(std::string *out, const int64_t *values, int n) {
    if (n == 0) { out->push_back(')'); return; }
    if (n > 3) out->push_back('(');
    for (int k = 0; k < n; ++k) pickle_int(out, values[k]);
    out->push_back(n == 1 ? (char)0x85 : n == 2 ? (char)0x86 : n == 3 ? (char)0x87 : 't');
}

// the pickle torch.save writes for a contiguous CPU int32 tensor (protocol 2)
std::string tensor_pickle_i32(const int64_t *shape, int ndim) {
    int64_t numel = 1, stride[4];
    for (int d = ndim - 1; d >= 0; --d) { stride[d] = numel; numel *= shape[d]; }
    std::string p;
    p += "\x80\x02" "ctorch._utils\n_rebuild_tensor_v2\nq";
    p.push_back('\0');
    p += "((X\x07";
    p.append(3, '\0');
    p += "storageq\x01" "ctorch\nIntStorage\nq\x02X\x01";
    p.append(3, '\0');
    p += "0q\x03X\x03";
    p.append(3, '\0');
    p += "cpuq\x04";
    pickle_int(&p, numel);
    p += "tq\x05Q";
    pickle_int(&p, 0);
    pickle_tuple(&p, shape, ndim);
    p += "q\x06";
    pickle_tuple(&p, stride, ndim);
    p += "q\x07\x89" "ccollections\nOrderedDict\nq\x08)Rq\ttq\nRq\x0b.";
    return p;
}

struct Record { std::string name; uint32_t crc, size, offset; };

void append_record(std::string *file, std::vector<Record> *records, const std::string &name,
                   const void *data, size_t bytes, size_t align) {
    Record record;
    record.name = name;
    record.size = (uint32_t)bytes;
    record.crc = crc32_of(static_cast<const unsigned char *>(data), bytes);
    record.offset = (uint32_t)file->size();
    // pad through the extra field so the payload starts `align`-byte aligned (torch aligns storages to 64)
    size_t extra = 0;
    const size_t start = file->size() + 30 + name.size();
    if (align > 1) {
        extra = (align - (start + 4) % align) % align + 4;         // 4-byte extra header + padding
    }
    put32(file, 0x04034b50u); put16(file, 20); put16(file, 0x0800); put16(file, 0);   // version, UTF-8 flag, stored
    put16(file, 0); put16(file, 0);                                                    // time, date
    put32(file, record.crc); put32(file, record.size); put32(file, record.size);
    put16(file, (uint16_t)name.size()); put16(file, (uint16_t)extra);
    *file += name;
    if (extra) {
        put16(file, 0x4246);                                        // "FB", the id torch's writer uses for padding
        put16(file, (uint16_t)(extra - 4));
        file->append(extra - 4, 'Z');
    }
    file->append(static_cast<const char *>(data), bytes);
    records->push_back(std::move(record));
}

int write_one_i32_unchecked(const char *path, const int32_t *data, const int64_t *shape, int ndim);

int write_one_i32(const char *path, const int32_t *data, const int64_t *shape, int ndim) {
    try {
        return write_one_i32_unchecked(path, data, shape, ndim);
    } catch (...) {
        return TORBI_B200_PT_IO_ERROR;
    }
}

int write_one_i32_unchecked(const char *path, const int32_t *data, const int64_t *shape, int ndim) {
    int64_t numel = 1;
    for (int d = 0; d < ndim; ++d) numel *= shape[d];
    if (ndim < 0 || ndim > 4 || numel < 0 || (uint64_t)numel * 4 > 0x7fffffffu) return TORBI_B200_PT_UNSUPPORTED;
    const std::string prefix = "archive/";
    const std::string pkl = tensor_pickle_i32(shape, ndim);
    std::string file;
    file.reserve((size_t)numel * 4 + 1024);
    std::vector<Record> records;
    append_record(&file, &records, prefix + "data.pkl", pkl.data(), pkl.size(), 1);
    append_record(&file, &records, prefix + "byteorder", "little", 6, 1);
    append_record(&file, &records, prefix + "data/0", data, (size_t)numel * 4, 64);
    append_record(&file, &records, prefix + "version", "3\n", 2, 1);
    const uint32_t cd_offset = (uint32_t)file.size();
    for (const Record &record : records) {
        put32(&file, 0x02014b50u); put16(&file, 20); put16(&file, 20); put16(&file, 0x0800); put16(&file, 0);
        put16(&file, 0); put16(&file, 0);
        put32(&file, record.crc); put32(&file, record.size); put32(&file, record.size);
        put16(&file, (uint16_t)record.name.size()); put16(&file, 0); put16(&file, 0);
        put16(&file, 0); put16(&file, 0); put32(&file, 0);
        put32(&file, record.offset);
        file += record.name;
    }
    const uint32_t cd_size = (uint32_t)file.size() - cd_offset;
    put32(&file, 0x06054b50u); put16(&file, 0); put16(&file, 0);
    put16(&file, (uint16_t)records.size()); put16(&file, (uint16_t)records.size());
    put32(&file, cd_size); put32(&file, cd_offset); put16(&file, 0);

    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (fd < 0) return TORBI_B200_PT_IO_ERROR;
    const char *p = file.data();
    size_t left = file.size();
    while (left > 0) {
        const ssize_t wrote = write(fd, p, left);
        if (wrote <= 0) { close(fd); return TORBI_B200_PT_IO_ERROR; }
        p += wrote;
        left -= (size_t)wrote;
    }
    return close(fd) == 0 ? TORBI_B200_PT_OK : TORBI_B200_PT_IO_ERROR;
}

}  // namespace

extern "C" {

int torbi_b200_pt_inspect(const char *const *paths, int64_t n, torbi_b200_pt_info *info, int threads) {
    if (!paths || !info || n < 0) return TORBI_B200_ERR_INVALID_ARGUMENT;
    try {
        std::atomic<int> failures{0};
        parallel_for(n, threads, [&](int64_t i) {
            if (inspect_one(paths[i], &info[i]) != TORBI_B200_PT_OK) failures.fetch_add(1);
        });
        return failures.load();
    } catch (...) {                      // (thread creation failure)
        return TORBI_B200_ERR_OUT_OF_MEMORY;
    }
}

int torbi_b200_pt_read(const char *const *paths, const torbi_b200_pt_info *info, int64_t n,
                       void *const *destinations, const int64_t *capacities, int threads) {
    if (!paths || !info || !destinations || !capacities || n < 0) return TORBI_B200_ERR_INVALID_ARGUMENT;
    try {
    // split large payloads so a batch of few big files still uses every thread
    struct Piece { int64_t file; uint64_t begin, end; };
    std::vector<Piece> pieces;
    const uint64_t kPiece = uint64_t(8) << 20;
    for (int64_t i = 0; i < n; ++i) {
        if (info[i].status != TORBI_B200_PT_OK || info[i].nbytes > capacities[i]) return TORBI_B200_ERR_INVALID_ARGUMENT;
        const uint64_t capacity = (uint64_t)capacities[i];
        for (uint64_t at = 0; at < capacity || at == 0; at += kPiece) {
            pieces.push_back({i, at, std::min(capacity, at + kPiece)});
            if (capacity == 0) break;
        }
    }
    std::atomic<int> failures{0};
    parallel_for((int64_t)pieces.size(), threads, [&](int64_t k) {
        const Piece &piece = pieces[k];
        const torbi_b200_pt_info &file = info[piece.file];
        char *dst = static_cast<char *>(destinations[piece.file]);
        const uint64_t payload_end = std::min<uint64_t>(piece.end, (uint64_t)file.nbytes);
        if (piece.begin < payload_end) {
            const int fd = open(paths[piece.file], O_RDONLY | O_CLOEXEC);
            if (fd < 0 || !read_at(fd, dst + piece.begin, payload_end - piece.begin, (uint64_t)file.data_offset + piece.begin))
                failures.fetch_add(1);
            if (fd >= 0) close(fd);
        }
        // zero padding behind the payload, like collate() (collate.py:26-31)
        const uint64_t zero_from = std::max<uint64_t>(piece.begin, (uint64_t)file.nbytes);
        if (zero_from < piece.end) std::memset(dst + zero_from, 0, piece.end - zero_from);
    });
    return failures.load();
    } catch (...) {
        return TORBI_B200_ERR_OUT_OF_MEMORY;
    }
}

int torbi_b200_pt_write_i32(const char *const *paths, const int32_t *const *data, const int64_t *shapes,
                            const int32_t *ndims, int64_t n, int threads) {
    if (!paths || !data || !shapes || !ndims || n < 0) return TORBI_B200_ERR_INVALID_ARGUMENT;
    try {
        std::atomic<int> failures{0};
        parallel_for(n, threads, [&](int64_t i) {
            if (write_one_i32(paths[i], data[i], shapes + 4 * i, ndims[i]) != TORBI_B200_PT_OK) failures.fetch_add(1);
        });
        return failures.load();
    } catch (...) {
        return TORBI_B200_ERR_OUT_OF_MEMORY;
    }
}

}  // extern "C"
