// See h5lite.h.  Format facts follow the public "HDF5 File Format Specification Version 2.0" (superblock
// v0, object header v1, B-tree v1, local heaps, symbol nodes, data layout v3, filter pipeline v1/v2).
#include "h5lite.h"

#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <fstream>

namespace h5l {
namespace {

const uint64_t UNDEF = ~0ull;
const unsigned char SIG[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};

struct Reader {
    std::vector<uint8_t> b;

    template <typename T> T rd(uint64_t off) const {
        if (off + sizeof(T) > b.size()) throw std::string("h5lite: read past end of file");
        T v;
        memcpy(&v, b.data() + off, sizeof(T));
        return v;
    }
    bool tag(uint64_t off, const char* t) const { return off + 4 <= b.size() && memcmp(b.data() + off, t, 4) == 0; }

    struct Msg { uint16_t type; uint8_t flags; uint64_t off; uint16_t size; };

    std::vector<Msg> messages(uint64_t addr) const {
        uint8_t ver = rd<uint8_t>(addr);
        if (ver != 1) throw std::string("h5lite: only version-1 object headers are supported");
        uint16_t nmsg = rd<uint16_t>(addr + 2);
        uint32_t hsize = rd<uint32_t>(addr + 8);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hsize}};
        std::vector<Msg> out;
        for (size_t bi = 0; bi < blocks.size() && out.size() < nmsg; ++bi) {
            uint64_t p = blocks[bi].first, end = p + blocks[bi].second;
            while (p + 8 <= end && out.size() < nmsg) {
                Msg m{rd<uint16_t>(p), rd<uint8_t>(p + 4), p + 8, rd<uint16_t>(p + 2)};
                p += 8 + m.size;
                if (m.type == 0x10) blocks.push_back({rd<uint64_t>(m.off), rd<uint64_t>(m.off + 8)});
                out.push_back(m);
            }
        }
        return out;
    }

    DType parse_dtype(uint64_t off) const {
        uint8_t cv = rd<uint8_t>(off), bf0 = rd<uint8_t>(off + 1);
        uint32_t size = rd<uint32_t>(off + 4);
        DType d;
        d.size = size;
        switch (cv & 0xF) {
            case 0:
                if (bf0 & 1) throw std::string("h5lite: big-endian integers unsupported");
                d.kind = (bf0 & 8) ? Kind::Int : Kind::UInt;
                break;
            case 1:
                if (bf0 & 1) throw std::string("h5lite: big-endian floats unsupported");
                if (size != 4 && size != 8) throw std::string("h5lite: unsupported float width");
                d.kind = Kind::Float;
                break;
            case 3: d.kind = Kind::String; break;
            default: throw std::string("h5lite: unsupported datatype class ") + std::to_string(cv & 0xF);
        }
        return d;
    }

    // returns false for a null dataspace
    bool parse_space(uint64_t off, Array& a) const {
        uint8_t ver = rd<uint8_t>(off), rank = rd<uint8_t>(off + 1), flags = rd<uint8_t>(off + 2);
        uint64_t p;
        if (ver == 1) p = off + 8;
        else if (ver == 2) { p = off + 4; if (rd<uint8_t>(off + 3) == 2) return false; }
        else throw std::string("h5lite: unsupported dataspace version");
        a.scalar = (rank == 0);
        a.dims.resize(rank);
        for (int i = 0; i < rank; ++i) a.dims[i] = rd<uint64_t>(p + 8 * i);
        if (flags & 1) {
            a.maxdims.resize(rank);
            for (int i = 0; i < rank; ++i) a.maxdims[i] = rd<uint64_t>(p + 8 * (rank + i));
        }
        return true;
    }

    void parse_attr(const Msg& m, std::map<std::string, Array>& attrs) const {
        uint64_t o = m.off;
        uint8_t ver = rd<uint8_t>(o);
        uint16_t nsz = rd<uint16_t>(o + 2), dsz = rd<uint16_t>(o + 4), ssz = rd<uint16_t>(o + 6);
        uint64_t p = o + (ver == 3 ? 9 : 8);
        auto pad = [&](uint64_t n) { return ver == 1 ? ((n + 7) & ~7ull) : n; };
        std::string name((const char*)b.data() + p, strnlen((const char*)b.data() + p, nsz));
        p += pad(nsz);
        Array a;
        a.dt = parse_dtype(p);
        p += pad(dsz);
        bool ok = parse_space(p, a);
        p += pad(ssz);
        if (!ok) return;
        uint64_t nbytes = a.count() * a.dt.size;
        if (p + nbytes > b.size()) throw std::string("h5lite: attribute data past end of file");
        a.raw.assign(b.begin() + p, b.begin() + p + nbytes);
        attrs[name] = std::move(a);
    }

    struct Filter { uint16_t id; std::vector<uint32_t> cd; };

    std::vector<Filter> parse_filters(uint64_t o) const {
        uint8_t ver = rd<uint8_t>(o), nf = rd<uint8_t>(o + 1);
        uint64_t p = o + (ver == 1 ? 8 : 2);
        std::vector<Filter> out;
        for (int i = 0; i < nf; ++i) {
            Filter f;
            f.id = rd<uint16_t>(p); p += 2;
            uint16_t nlen = 0;
            if (ver == 1 || f.id >= 256) { nlen = rd<uint16_t>(p); p += 2; }
            p += 2;  // flags
            uint16_t ncd = rd<uint16_t>(p); p += 2;
            p += (ver == 1) ? ((nlen + 7) & ~7u) : nlen;
            for (int k = 0; k < ncd; ++k) { f.cd.push_back(rd<uint32_t>(p)); p += 4; }
            if (ver == 1 && (ncd & 1)) p += 4;
            out.push_back(f);
        }
        return out;
    }

    struct Chunk { std::vector<uint64_t> offs; uint32_t size, mask; uint64_t addr; };

    void chunk_btree(uint64_t addr, int rank1, std::vector<Chunk>& out) const {
        if (!tag(addr, "TREE")) throw std::string("h5lite: bad chunk B-tree node");
        uint8_t level = rd<uint8_t>(addr + 5);
        uint16_t nent = rd<uint16_t>(addr + 6);
        uint64_t p = addr + 24, keysz = 8 + 8ull * rank1;
        for (int i = 0; i < nent; ++i) {
            Chunk c;
            c.size = rd<uint32_t>(p);
            c.mask = rd<uint32_t>(p + 4);
            for (int k = 0; k < rank1 - 1; ++k) c.offs.push_back(rd<uint64_t>(p + 8 + 8 * k));
            c.addr = rd<uint64_t>(p + keysz);
            p += keysz + 8;
            if (level > 0) chunk_btree(c.addr, rank1, out);
            else out.push_back(c);
        }
    }

    static std::vector<uint8_t> inflate_all(const uint8_t* src, size_t n, size_t hint) {
        std::vector<uint8_t> out(std::max<size_t>(hint + 64, 256));
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) throw std::string("h5lite: zlib init failed");
        zs.next_in = const_cast<Bytef*>(src);
        zs.avail_in = (uInt)n;
        size_t have = 0;
        for (;;) {
            zs.next_out = out.data() + have;
            zs.avail_out = (uInt)(out.size() - have);
            int rc = inflate(&zs, Z_NO_FLUSH);
            have = out.size() - zs.avail_out;
            if (rc == Z_STREAM_END) break;
            if (rc != Z_OK) { inflateEnd(&zs); throw std::string("h5lite: zlib inflate failed"); }
            if (zs.avail_out == 0) out.resize(out.size() * 2);
        }
        inflateEnd(&zs);
        out.resize(have);
        return out;
    }

    void read_dataset(const std::vector<Msg>& msgs, Node& n) const {
        Array& a = n.data;
        const Msg* layout = nullptr;
        std::vector<Filter> filters;
        bool have_space = false;
        for (auto& m : msgs) {
            if (m.type == 0x1) have_space = parse_space(m.off, a);
            else if (m.type == 0x3) a.dt = parse_dtype(m.off);
            else if (m.type == 0x8) layout = &m;
            else if (m.type == 0xB) filters = parse_filters(m.off);
            else if (m.type == 0xC) parse_attr(m, n.attrs);
        }
        if (!layout || !have_space) throw std::string("h5lite: dataset without layout/dataspace");
        uint64_t o = layout->off;
        if (rd<uint8_t>(o) != 3) throw std::string("h5lite: only data layout message v3 is supported");
        uint8_t cls = rd<uint8_t>(o + 1);
        uint64_t nbytes = a.count() * a.dt.size;
        a.raw.assign(nbytes, 0);
        if (cls == 0) {
            uint16_t sz = rd<uint16_t>(o + 2);
            memcpy(a.raw.data(), b.data() + o + 4, std::min<uint64_t>(sz, nbytes));
        } else if (cls == 1) {
            uint64_t addr = rd<uint64_t>(o + 2);
            if (addr != UNDEF && nbytes) {
                if (addr + nbytes > b.size()) throw std::string("h5lite: dataset past end of file");
                memcpy(a.raw.data(), b.data() + addr, nbytes);
            }
        } else if (cls == 2) {
            int rank1 = rd<uint8_t>(o + 2);
            int rank = rank1 - 1;
            if (rank != (int)a.dims.size()) throw std::string("h5lite: chunk rank mismatch");
            uint64_t baddr = rd<uint64_t>(o + 3);
            std::vector<uint64_t> cd(rank);
            uint64_t celems = 1;
            for (int i = 0; i < rank; ++i) { cd[i] = rd<uint32_t>(o + 11 + 4 * i); celems *= cd[i]; }
            if (baddr == UNDEF || nbytes == 0) return;
            std::vector<Chunk> chunks;
            chunk_btree(baddr, rank1, chunks);
            const uint32_t es = a.dt.size;
            for (auto& c : chunks) {
                if (c.addr + c.size > b.size()) throw std::string("h5lite: chunk past end of file");
                std::vector<uint8_t> raw(b.begin() + c.addr, b.begin() + c.addr + c.size);
                for (int fi = (int)filters.size() - 1; fi >= 0; --fi) {
                    if (c.mask & (1u << fi)) continue;
                    switch (filters[fi].id) {
                        case 1: raw = inflate_all(raw.data(), raw.size(), celems * es + 4); break;
                        case 2: {  // shuffle
                            size_t ne = raw.size() / es;
                            std::vector<uint8_t> t(raw);
                            for (size_t e = 0; e < ne; ++e)
                                for (uint32_t k = 0; k < es; ++k) t[e * es + k] = raw[k * ne + e];
                            raw.swap(t);
                            break;
                        }
                        case 3: if (raw.size() >= 4) raw.resize(raw.size() - 4); break;  // fletcher32 (not verified)
                        default: throw std::string("h5lite: unsupported filter id ") + std::to_string(filters[fi].id);
                    }
                }
                if (raw.size() < celems * es) throw std::string("h5lite: short chunk");
                // copy the chunk's valid region into the dataset, row by row over the last dimension
                std::vector<uint64_t> idx(rank, 0);
                if (rank == 0) continue;
                uint64_t last = rank - 1;
                uint64_t run = (c.offs[last] < a.dims[last]) ? std::min(cd[last], a.dims[last] - c.offs[last]) : 0;
                if (!run) continue;
                for (;;) {
                    bool inside = true;
                    uint64_t dst = 0, src = 0;
                    for (int d = 0; d < rank; ++d) {
                        uint64_t g = c.offs[d] + idx[d];
                        if (g >= a.dims[d]) inside = false;
                        dst = dst * a.dims[d] + g;
                        src = src * cd[d] + idx[d];
                    }
                    if (inside) memcpy(a.raw.data() + dst * es, raw.data() + src * es, run * es);
                    int d = rank - 2;
                    for (; d >= 0; --d) { if (++idx[d] < cd[d]) break; idx[d] = 0; }
                    if (d < 0) break;
                }
            }
        } else {
            throw std::string("h5lite: unknown layout class");
        }
    }

    void group_entries(uint64_t addr, uint64_t heap_data, std::vector<std::pair<std::string, uint64_t>>& out) const {
        if (tag(addr, "TREE")) {
            uint16_t nent = rd<uint16_t>(addr + 6);
            uint64_t p = addr + 24;
            for (int i = 0; i < nent; ++i) { group_entries(rd<uint64_t>(p + 8), heap_data, out); p += 16; }
        } else if (tag(addr, "SNOD")) {
            uint16_t nsym = rd<uint16_t>(addr + 6);
            uint64_t p = addr + 8;
            for (int i = 0; i < nsym; ++i) {
                uint64_t s = heap_data + rd<uint64_t>(p);
                if (s >= b.size()) throw std::string("h5lite: bad link name offset");
                out.push_back({std::string((const char*)b.data() + s, strnlen((const char*)b.data() + s, b.size() - s)),
                               rd<uint64_t>(p + 8)});
                p += 40;
            }
        } else {
            throw std::string("h5lite: bad group B-tree/SNOD node");
        }
    }

    std::unique_ptr<Node> read_object(uint64_t addr, int depth = 0) const {
        if (depth > 64) throw std::string("h5lite: group nesting too deep");
        auto msgs = messages(addr);
        std::unique_ptr<Node> n(new Node);
        bool is_dset = false, is_group = false;
        for (auto& m : msgs) { if (m.type == 0x8) is_dset = true; if (m.type == 0x11) is_group = true; }
        if (is_group) {
            for (auto& m : msgs) {
                if (m.type == 0x11) {
                    uint64_t bt = rd<uint64_t>(m.off), hp = rd<uint64_t>(m.off + 8);
                    if (!tag(hp, "HEAP")) throw std::string("h5lite: bad local heap");
                    std::vector<std::pair<std::string, uint64_t>> ents;
                    group_entries(bt, rd<uint64_t>(hp + 24), ents);
                    for (auto& e : ents) n->children[e.first] = read_object(e.second, depth + 1);
                } else if (m.type == 0xC) parse_attr(m, n->attrs);
            }
        } else if (is_dset) {
            n->is_group = false;
            read_dataset(msgs, *n);
        } else {
            for (auto& m : msgs) if (m.type == 0x2 || m.type == 0x6)
                throw std::string("h5lite: new-style (link-message) groups are not supported");
        }
        return n;
    }
};

// ---------------------------------------------------------------------------------------------- writer
struct Writer {
    std::vector<uint8_t> buf;
    static const int K_LEAF = 4, K_INT = 16;

    template <typename T> static void put(std::vector<uint8_t>& v, T x) {
        size_t n = v.size(); v.resize(n + sizeof(T)); memcpy(v.data() + n, &x, sizeof(T));
    }
    static void pad8(std::vector<uint8_t>& v) { while (v.size() % 8) v.push_back(0); }

    uint64_t alloc(const std::vector<uint8_t>& d) {
        while (buf.size() % 8) buf.push_back(0);
        uint64_t a = buf.size();
        buf.insert(buf.end(), d.begin(), d.end());
        return a;
    }

    static std::vector<uint8_t> dtype_msg(const DType& d) {
        std::vector<uint8_t> m;
        if (d.kind == Kind::Int || d.kind == Kind::UInt) {
            m = {0x10, (uint8_t)(d.kind == Kind::Int ? 8 : 0), 0, 0};
            put<uint32_t>(m, d.size); put<uint16_t>(m, 0); put<uint16_t>(m, 8 * d.size);
        } else if (d.kind == Kind::Float) {
            m = {0x11, 0x20, (uint8_t)(d.size == 4 ? 31 : 63), 0};
            put<uint32_t>(m, d.size); put<uint16_t>(m, 0); put<uint16_t>(m, 8 * d.size);
            if (d.size == 4) { m.insert(m.end(), {23, 8, 0, 23}); put<uint32_t>(m, 127); }
            else { m.insert(m.end(), {52, 11, 0, 52}); put<uint32_t>(m, 1023); }
        } else {
            m = {0x13, 0, 0, 0};
            put<uint32_t>(m, std::max<uint32_t>(d.size, 1));
        }
        return m;
    }
    static std::vector<uint8_t> space_msg(const Array& a) {
        std::vector<uint8_t> m{1, (uint8_t)(a.scalar ? 0 : a.dims.size()), 0, 0, 0, 0, 0, 0};
        if (!a.scalar) for (auto d : a.dims) put<uint64_t>(m, d);
        return m;
    }
    static std::vector<uint8_t> attr_msg(const std::string& name, const Array& a) {
        auto dt = dtype_msg(a.dt), sp = space_msg(a);
        std::vector<uint8_t> m{1, 0};
        put<uint16_t>(m, name.size() + 1); put<uint16_t>(m, dt.size()); put<uint16_t>(m, sp.size());
        m.insert(m.end(), name.begin(), name.end()); m.push_back(0); pad8(m);
        m.insert(m.end(), dt.begin(), dt.end()); pad8(m);
        m.insert(m.end(), sp.begin(), sp.end()); pad8(m);
        m.insert(m.end(), a.raw.begin(), a.raw.end());
        return m;
    }
    uint64_t object_header(const std::vector<std::pair<uint16_t, std::vector<uint8_t>>>& msgs) {
        std::vector<uint8_t> body;
        for (auto& m : msgs) {
            std::vector<uint8_t> mb = m.second; pad8(mb);
            if (mb.size() > 65535) throw std::string("h5lite: header message too large (attribute > 64 KiB?)");
            put<uint16_t>(body, m.first); put<uint16_t>(body, mb.size()); put<uint32_t>(body, 0);
            body.insert(body.end(), mb.begin(), mb.end());
        }
        std::vector<uint8_t> h{1, 0};
        put<uint16_t>(h, msgs.size()); put<uint32_t>(h, 1); put<uint32_t>(h, body.size()); put<uint32_t>(h, 0);
        h.insert(h.end(), body.begin(), body.end());
        return alloc(h);
    }
    uint64_t write_dataset(const Node& n) {
        std::vector<std::pair<uint16_t, std::vector<uint8_t>>> msgs;
        msgs.push_back({0x1, space_msg(n.data)});
        msgs.push_back({0x3, dtype_msg(n.data.dt)});
        msgs.push_back({0x5, {2, 2, 2, 0}});
        uint64_t addr = n.data.raw.empty() ? UNDEF : alloc(n.data.raw);
        std::vector<uint8_t> lay{3, 1};
        put<uint64_t>(lay, addr); put<uint64_t>(lay, n.data.raw.size());
        msgs.push_back({0x8, lay});
        for (auto& kv : n.attrs) msgs.push_back({0xC, attr_msg(kv.first, kv.second)});
        return object_header(msgs);
    }
    uint64_t write_group(const Node& g, uint64_t* bt_out = nullptr, uint64_t* heap_out = nullptr) {
        std::vector<std::pair<std::string, uint64_t>> ents;
        for (auto& kv : g.children)
            ents.push_back({kv.first, kv.second->is_group ? write_group(*kv.second) : write_dataset(*kv.second)});
        std::vector<uint8_t> heap(8, 0);
        std::vector<uint64_t> noff;
        for (auto& e : ents) {
            noff.push_back(heap.size());
            heap.insert(heap.end(), e.first.begin(), e.first.end()); heap.push_back(0); pad8(heap);
        }
        uint64_t free_off = heap.size();
        put<uint64_t>(heap, 1); put<uint64_t>(heap, 16);
        uint64_t heap_data = alloc(heap);
        std::vector<uint8_t> hh{'H', 'E', 'A', 'P', 0, 0, 0, 0};
        put<uint64_t>(hh, heap.size()); put<uint64_t>(hh, free_off); put<uint64_t>(hh, heap_data);
        uint64_t heap_addr = alloc(hh);
        const size_t per = 2 * K_LEAF;
        std::vector<std::pair<uint64_t, uint64_t>> snods;
        for (size_t i = 0; i < std::max<size_t>(ents.size(), 1); i += per) {
            size_t cnt = std::min(per, ents.size() - std::min(ents.size(), i));
            std::vector<uint8_t> sn{'S', 'N', 'O', 'D', 1, 0};
            put<uint16_t>(sn, cnt);
            for (size_t k = i; k < i + cnt; ++k) {
                put<uint64_t>(sn, noff[k]); put<uint64_t>(sn, ents[k].second); put<uint32_t>(sn, 0); put<uint32_t>(sn, 0);
                sn.resize(sn.size() + 16, 0);
            }
            sn.resize(8 + 40 * per, 0);
            snods.push_back({alloc(sn), cnt ? noff[i + cnt - 1] : 0});
        }
        if (snods.size() > 2 * K_INT) throw std::string("h5lite: too many children in one group");
        std::vector<uint8_t> bt{'T', 'R', 'E', 'E', 0, 0};
        put<uint16_t>(bt, snods.size()); put<uint64_t>(bt, UNDEF); put<uint64_t>(bt, UNDEF); put<uint64_t>(bt, 0);
        for (auto& s : snods) { put<uint64_t>(bt, s.first); put<uint64_t>(bt, s.second); }
        bt.resize(24 + 8 + 2 * K_INT * 16, 0);
        uint64_t bt_addr = alloc(bt);
        std::vector<std::pair<uint16_t, std::vector<uint8_t>>> msgs;
        std::vector<uint8_t> st;
        put<uint64_t>(st, bt_addr); put<uint64_t>(st, heap_addr);
        msgs.push_back({0x11, st});
        for (auto& kv : g.attrs) msgs.push_back({0xC, attr_msg(kv.first, kv.second)});
        if (bt_out) *bt_out = bt_addr;
        if (heap_out) *heap_out = heap_addr;
        return object_header(msgs);
    }
};

template <typename T, typename S> void conv_loop(const uint8_t* raw, uint64_t n, std::vector<T>& out) {
    for (uint64_t i = 0; i < n; ++i) { S v; memcpy(&v, raw + i * sizeof(S), sizeof(S)); out[i] = (T)v; }
}

}  // namespace

std::unique_ptr<Node> load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::string("h5lite: cannot open '") + path + "'";
    Reader r;
    r.b.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    if (r.b.size() < 96 || memcmp(r.b.data(), SIG, 8)) throw std::string("h5lite: '") + path + "' is not an HDF5 file";
    uint8_t ver = r.b[8];
    if (ver > 1) throw std::string("h5lite: only superblock v0/v1 supported");
    if (r.b[13] != 8 || r.b[14] != 8) throw std::string("h5lite: only 8-byte offsets/lengths supported");
    uint64_t off = 24 + (ver == 1 ? 4 : 0) + 32;
    uint64_t root_ohdr = r.rd<uint64_t>(off + 8);
    return r.read_object(root_ohdr);
}

void save(const Node& root, const std::string& path) {
    Writer w;
    w.buf.assign(96, 0);
    uint64_t bt = 0, heap = 0;
    uint64_t oh = w.write_group(root, &bt, &heap);
    std::vector<uint8_t> sb(SIG, SIG + 8);
    sb.insert(sb.end(), {0, 0, 0, 0, 0, 8, 8, 0});
    Writer::put<uint16_t>(sb, Writer::K_LEAF); Writer::put<uint16_t>(sb, Writer::K_INT); Writer::put<uint32_t>(sb, 0);
    Writer::put<uint64_t>(sb, 0); Writer::put<uint64_t>(sb, UNDEF); Writer::put<uint64_t>(sb, w.buf.size()); Writer::put<uint64_t>(sb, UNDEF);
    Writer::put<uint64_t>(sb, 0); Writer::put<uint64_t>(sb, oh); Writer::put<uint32_t>(sb, 1); Writer::put<uint32_t>(sb, 0);
    Writer::put<uint64_t>(sb, bt); Writer::put<uint64_t>(sb, heap);
    memcpy(w.buf.data(), sb.data(), 96);
    std::string tmp = path + ".tmp~";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) throw std::string("h5lite: cannot write '") + path + "'";
        f.write((const char*)w.buf.data(), w.buf.size());
        if (!f) throw std::string("h5lite: write failed for '") + path + "'";
    }
    if (rename(tmp.c_str(), path.c_str())) throw std::string("h5lite: rename failed for '") + path + "'";
}

Node* find(Node* base, const std::string& path) {
    Node* n = base;
    size_t i = 0;
    while (n && i < path.size()) {
        size_t j = path.find('/', i);
        if (j == std::string::npos) j = path.size();
        std::string part = path.substr(i, j - i);
        i = j + 1;
        if (part.empty() || part == ".") continue;
        if (!n->is_group) return nullptr;
        auto it = n->children.find(part);
        if (it == n->children.end()) return nullptr;
        n = it->second.get();
    }
    return n;
}

Node* ensure_group(Node* base, const std::string& path) {
    Node* n = base;
    size_t i = 0;
    while (i < path.size()) {
        size_t j = path.find('/', i);
        if (j == std::string::npos) j = path.size();
        std::string part = path.substr(i, j - i);
        i = j + 1;
        if (part.empty() || part == ".") continue;
        auto& c = n->children[part];
        if (!c) c.reset(new Node);
        n = c.get();
    }
    return n;
}

template <typename T> std::vector<T> as(const Array& a) {
    uint64_t n = a.count();
    std::vector<T> out(n);
    if (a.raw.size() < n * a.dt.size) throw std::string("h5lite: array storage shorter than its shape");
    const uint8_t* r = a.raw.data();
    switch (a.dt.kind) {
        case Kind::Float:
            if (a.dt.size == 4) conv_loop<T, float>(r, n, out); else conv_loop<T, double>(r, n, out);
            break;
        case Kind::Int:
            switch (a.dt.size) {
                case 1: conv_loop<T, int8_t>(r, n, out); break;
                case 2: conv_loop<T, int16_t>(r, n, out); break;
                case 4: conv_loop<T, int32_t>(r, n, out); break;
                case 8: conv_loop<T, int64_t>(r, n, out); break;
                default: throw std::string("h5lite: bad integer width");
            }
            break;
        case Kind::UInt:
            switch (a.dt.size) {
                case 1: conv_loop<T, uint8_t>(r, n, out); break;
                case 2: conv_loop<T, uint16_t>(r, n, out); break;
                case 4: conv_loop<T, uint32_t>(r, n, out); break;
                case 8: conv_loop<T, uint64_t>(r, n, out); break;
                default: throw std::string("h5lite: bad integer width");
            }
            break;
        default: throw std::string("h5lite: cannot convert a string array to numbers");
    }
    return out;
}
template std::vector<float> as<float>(const Array&);
template std::vector<double> as<double>(const Array&);
template std::vector<int> as<int>(const Array&);
template std::vector<long> as<long>(const Array&);
template std::vector<unsigned> as<unsigned>(const Array&);

std::vector<std::string> as_strings(const Array& a) {
    if (a.dt.kind != Kind::String) throw std::string("h5lite: not a string array");
    std::vector<std::string> out;
    uint64_t n = a.count();
    for (uint64_t i = 0; i < n; ++i) {
        const char* p = (const char*)a.raw.data() + i * a.dt.size;
        out.emplace_back(p, strnlen(p, a.dt.size));
    }
    return out;
}

template <typename T> struct KindOf;
template <> struct KindOf<float> { static const Kind k = Kind::Float; };
template <> struct KindOf<double> { static const Kind k = Kind::Float; };
template <> struct KindOf<int> { static const Kind k = Kind::Int; };
template <> struct KindOf<long> { static const Kind k = Kind::Int; };

template <typename T> Array make_array(const std::vector<T>& v, const std::vector<uint64_t>& dims) {
    Array a;
    a.dt.kind = KindOf<T>::k;
    a.dt.size = sizeof(T);
    a.dims = dims;
    a.scalar = dims.empty();
    if (a.count() != v.size()) throw std::string("h5lite: make_array shape mismatch");
    a.raw.resize(v.size() * sizeof(T));
    if (!v.empty()) memcpy(a.raw.data(), v.data(), a.raw.size());
    return a;
}
template Array make_array<float>(const std::vector<float>&, const std::vector<uint64_t>&);
template Array make_array<double>(const std::vector<double>&, const std::vector<uint64_t>&);
template Array make_array<int>(const std::vector<int>&, const std::vector<uint64_t>&);
template Array make_array<long>(const std::vector<long>&, const std::vector<uint64_t>&);

Array make_string_array(const std::vector<std::string>& v) {
    Array a;
    a.dt.kind = Kind::String;
    size_t w = 1;
    for (auto& s : v) w = std::max(w, s.size());
    a.dt.size = w;
    a.dims = {v.size()};
    a.raw.assign(v.size() * w, 0);
    for (size_t i = 0; i < v.size(); ++i) memcpy(a.raw.data() + i * w, v[i].data(), v[i].size());
    return a;
}
Array make_string_scalar(const std::string& s) {
    Array a;
    a.dt.kind = Kind::String;
    a.dt.size = s.size() + 1;
    a.scalar = true;
    a.raw.assign(s.size() + 1, 0);
    memcpy(a.raw.data(), s.data(), s.size());
    return a;
}

}  // namespace h5l
