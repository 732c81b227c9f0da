// reading OpenFHE's serialized evaluation keys straight into the engine's staging buffers.
//
// The reference serialises keys with cereal's PORTABLE BINARY archive (core/include/utils/serial.h:99-115;
// examples/boolean-serial-binary.cpp:76-88: Serial::SerializeToFile(path, cc.GetRefreshKey() / cc.GetSwitchKey(),
// SerType::BINARY)) and a user gets them back as a tree of shared_ptr<RingGSWEvalKeyImpl> / NativePoly / NativeVector
// objects (boolean-serial-binary.cpp:115-131) that GPUSetup then walks and copies coefficient by coefficient
// (bootstrapping.cu:933-975).  Here the byte stream is INDEXED once (no OpenFHE object is ever built): the parser
// below understands exactly the object graph these two key types produce and records where every polynomial / key
// switching row lives in the stream; the setup path then gathers the raw little-endian coefficients from those offsets
// into its pinned staging chunks and uploads them -- the same chunks a flat key array would have produced.
//
// Stream grammar (cereal 1.3 PortableBinaryOutputArchive as used by the reference; all integers little-endian):
//   archive      := u8 endian(=1) object
//   shared_ptr<T>:= u32 polymorphic_id(=0x40000000: static type == dynamic type) u32 id [T]      (T follows iff id has
//                   bit 31 set, i.e. the first time this pointer is written; cereal/types/memory.hpp, polymorphic.hpp)
//   unique_ptr<T>:= u32 polymorphic_id(=0x40000000) u8 valid [T]
//   class T      := [u32 version] fields      (the version is written ONCE per type and archive: cereal class versioning)
//   vector<X>    := u64 size X*
//   RingGSWACCKeyImpl   := vector<vector<vector<shared_ptr<RingGSWEvalKeyImpl>>>>                   rgsw-acckey.h:157-159
//   RingGSWEvalKeyImpl  := vector<vector<NativePoly>>                                               rgsw-evalkey.h:149-151
//   NativePoly (PolyImpl<NativeVector>) := unique_ptr<NativeVector> u32 format shared_ptr<ILNativeParams>
//   NativeVector        := u64 size, size x u64 raw, NativeInteger modulus                          mubintvecnat.h:586-595
//   NativeInteger       := u64                                                                      ubintnat.h:1997-2001
//   ILNativeParams      := [u32 version ElemParams] u32 cyclotomicOrder u32 ringDimension u8 isPowerOfTwo
//                          u64 ciphertextModulus u64 rootOfUnity u64 bigCiphertextModulus u64 bigRootOfUnity
//   LWESwitchingKeyImpl := vector<vector<vector<NativeVector>>> A, vector<vector<vector<NativeInteger>>> B
//                          (lwe-keyswitchkey.h:104-107; vector<NativeInteger> = u64 size + raw u64 values,
//                          mubintvecnat.h:668-676)
#include <cstring>
#include <string>
#include <vector>

#include "engine.cuh"

namespace tfhe_b200 {

namespace {

struct Cursor {
    const unsigned char* p;
    size_t n, pos = 0;
    bool ok = true;
    std::string err;

    bool fail(const std::string& m) {
        if (ok) {
            ok = false;
            err = m + " (at byte " + std::to_string(pos) + ")";
        }
        return false;
    }
    bool need(size_t k) { return (ok && k <= n - pos) ? true : fail("serialized stream is truncated"); }
    // `count` 8-byte words (no overflow for hostile counts)
    bool need_words(u64 count) { return (ok && count <= (n - pos) / 8) ? true : fail("serialized stream is truncated"); }
    u32 get_u8() {
        if (!need(1))
            return 0;
        return p[pos++];
    }
    u32 get_u32() {
        if (!need(4))
            return 0;
        u32 v;
        memcpy(&v, p + pos, 4);
        pos += 4;
        return v;
    }
    u64 get_u64() {
        if (!need(8))
            return 0;
        u64 v;
        memcpy(&v, p + pos, 8);
        pos += 8;
        return v;
    }
    // class version: present the first time a type occurs in the archive
    void version(bool& seen) {
        if (!seen) {
            seen = true;
            get_u32();
        }
    }
    // shared_ptr header; returns true when the pointee follows
    bool shared_ptr_new() {
        const u32 pid = get_u32();
        if (ok && pid != 0x40000000u)
            fail("unexpected polymorphic id (a derived type or a null pointer where a key object was expected)");
        const u32 id = get_u32();
        return ok && (id & 0x80000000u);
    }
};

}  // namespace

int index_serialized_acc_key(const void* data, size_t bytes, SerializedAccKey* out, std::string* err) {
    Cursor c{(const unsigned char*)data, bytes};
    SerializedAccKey& k = *out;
    k = SerializedAccKey();
    k.base = (const unsigned char*)data;
    bool v_acc = false, v_ek = false, v_poly = false, v_vec = false, v_int = false, v_ilp = false, v_ep = false;
    std::vector<size_t> offs;        // coefficient offsets of the polynomials actually present, stream order
    std::vector<char> ek_null;       // per evaluation key: null pointer?
    if (c.get_u8() != 1)
        c.fail("not a little-endian cereal portable-binary archive");
    if (!c.shared_ptr_new())
        c.fail("the refreshing key pointer is null");
    c.version(v_acc);
    k.dim[0] = c.get_u64();
    for (u64 a = 0; c.ok && a < k.dim[0]; a++) {
        const u64 d1 = c.get_u64();
        if (a == 0)
            k.dim[1] = d1;
        else if (d1 != k.dim[1])
            c.fail("ragged refreshing key");
        for (u64 b = 0; c.ok && b < d1; b++) {
            const u64 d2 = c.get_u64();
            if (a == 0 && b == 0)
                k.dim[2] = d2;
            else if (d2 != k.dim[2])
                c.fail("ragged refreshing key");
            if (c.ok && (d2 > (1u << 24) || k.dim[0] * k.dim[1] * d2 > (1ull << 28)))
                c.fail("implausible refreshing key dimensions");
            for (u64 e = 0; c.ok && e < d2; e++) {
                // DM keys hold null pointers for the refresh digit a0 = 0 (rgsw-acc-dm.cpp:64-66 never fills them); a null
                // polymorphic pointer is a single u32 0 (cereal/types/polymorphic.hpp)
                if (c.need(4)) {
                    u32 pid;
                    memcpy(&pid, c.p + c.pos, 4);
                    if (pid == 0) {
                        c.pos += 4;
                        ek_null.push_back(1);
                        continue;
                    }
                }
                ek_null.push_back(0);
                if (!c.shared_ptr_new()) {
                    c.fail("shared evaluation-key objects are not supported");
                    break;
                }
                c.version(v_ek);
                const u64 L = c.get_u64();
                if (c.ok && k.rows == 0)
                    k.rows = L;
                else if (L != k.rows)
                    c.fail("ragged evaluation key");
                if (c.ok && L > 64)
                    c.fail("implausible number of RGSW rows");
                for (u64 l = 0; c.ok && l < L; l++) {
                    if (c.get_u64() != 2)
                        c.fail("an RGSW row must hold two polynomials");
                    for (int j = 0; c.ok && j < 2; j++) {
                        c.version(v_poly);
                        if (c.get_u32() != 0x40000000u || c.get_u8() != 1)
                            c.fail("polynomial without coefficient vector");
                        c.version(v_vec);
                        const u64 len = c.get_u64();
                        if (c.ok && k.N == 0)
                            k.N = len;
                        else if (len != k.N)
                            c.fail("polynomials of different length");
                        if (!c.need_words(len))
                            break;
                        offs.push_back(c.pos);
                        c.pos += len * 8;
                        c.version(v_int);
                        const u64 mod = c.get_u64();
                        if (c.ok && k.Q == 0)
                            k.Q = mod;
                        else if (mod != k.Q)
                            c.fail("polynomials with different moduli");
                        if (c.get_u32() != 0)
                            c.fail("bootstrapping-key polynomials must be in EVALUATION format");
                        if (c.shared_ptr_new()) {   // ring parameters: written once, shared by every polynomial
                            c.version(v_ilp);
                            c.version(v_ep);
                            c.get_u32();                      // cyclotomic order
                            c.get_u32();                      // ring dimension
                            c.get_u8();                       // isPowerOfTwo
                            c.get_u64();                      // ciphertext modulus
                            k.psi = c.get_u64();              // root of unity the polynomials were transformed with
                            c.get_u64();
                            c.get_u64();
                        }
                    }
                }
            }
        }
    }
    if (c.ok && c.pos != c.n)
        c.fail("trailing bytes after the refreshing key");
    if (c.ok && (k.rows == 0 || k.N == 0))
        c.fail("the refreshing key holds no polynomials");
    if (!c.ok) {
        if (err)
            *err = "refreshing key: " + c.err;
        return TFHE_B200_EINVAL;
    }
    // flat polynomial index: null evaluation keys occupy their slots (they read as zero)
    const size_t per = (size_t)k.rows * 2;
    k.poly_off.reserve(ek_null.size() * per);
    size_t next = 0;
    for (char isnull : ek_null)
        for (size_t x = 0; x < per; x++)
            k.poly_off.push_back(isnull ? (size_t)-1 : offs[next++]);
    return 0;
}

int index_serialized_switch_key(const void* data, size_t bytes, SerializedSwitchKey* out, std::string* err) {
    Cursor c{(const unsigned char*)data, bytes};
    SerializedSwitchKey& k = *out;
    k = SerializedSwitchKey();
    k.base = (const unsigned char*)data;
    bool v_key = false, v_vec = false, v_int = false;
    if (c.get_u8() != 1)
        c.fail("not a little-endian cereal portable-binary archive");
    if (!c.shared_ptr_new())
        c.fail("the switching key pointer is null");
    c.version(v_key);
    // m_keyA [N][baseKS][dKS] NativeVector(n)
    k.N = c.get_u64();
    if (c.ok && k.N > (1u << 16))
        c.fail("implausible switching key dimensions");
    for (u64 i = 0; c.ok && i < k.N; i++) {
        const u64 bks = c.get_u64();
        if (i == 0)
            k.baseKS = bks;
        else if (bks != k.baseKS)
            c.fail("ragged switching key");
        if (c.ok && bks > (1u << 16))
            c.fail("implausible switching key dimensions");
        for (u64 a = 0; c.ok && a < bks; a++) {
            const u64 dks = c.get_u64();
            if (i == 0 && a == 0) {
                k.dKS = dks;
                if (c.ok && dks <= 64 && k.N * k.baseKS * dks <= bytes / 8)   // a row takes at least 8 bytes of stream
                    k.rowA_off.reserve((size_t)(k.N * k.baseKS * dks));
            }
            else if (dks != k.dKS)
                c.fail("ragged switching key");
            if (c.ok && dks > 64)
                c.fail("implausible switching key dimensions");
            for (u64 j = 0; c.ok && j < dks; j++) {
                c.version(v_vec);
                const u64 len = c.get_u64();
                if (c.ok && k.n == 0)
                    k.n = len;
                else if (len != k.n)
                    c.fail("key switching rows of different length");
                if (!c.need_words(len))
                    break;
                k.rowA_off.push_back(c.pos);
                c.pos += len * 8;
                c.version(v_int);
                const u64 mod = c.get_u64();
                if (c.ok && k.qKS == 0)
                    k.qKS = mod;
                else if (mod != k.qKS)
                    c.fail("key switching rows with different moduli");
            }
        }
    }
    // m_keyB [N][baseKS] vector<NativeInteger>(dKS) = u64 size + raw values
    if (c.get_u64() != k.N)
        c.fail("switching key: A and B disagree");
    for (u64 i = 0; c.ok && i < k.N; i++) {
        if (c.get_u64() != k.baseKS)
            c.fail("switching key: A and B disagree");
        for (u64 a = 0; c.ok && a < k.baseKS; a++) {
            if (c.get_u64() != k.dKS)
                c.fail("switching key: A and B disagree");
            if (!c.need_words(k.dKS))
                break;
            k.rowB_off.push_back(c.pos);
            c.pos += k.dKS * 8;
        }
    }
    if (c.ok && c.pos != c.n)
        c.fail("trailing bytes after the switching key");
    if (!c.ok) {
        if (err)
            *err = "switching key: " + c.err;
        return TFHE_B200_EINVAL;
    }
    return 0;
}

// words [off, off + cnt) of the flat bootstrapping key (element order of tfhe_b200_setup: consecutive polynomials)
void SerializedAccKey::gather(u64* dst, size_t off, size_t cnt) const {
    while (cnt) {
        const size_t poly = off / N, k0 = off % N, take = std::min<size_t>(cnt, N - k0);
        const size_t src = poly_off[poly];
        if (src == (size_t)-1)
            memset(dst, 0, take * 8);                 // DM: the never-read a0 = 0 entries
        else
            memcpy(dst, base + src + k0 * 8, take * 8);
        dst += take;
        off += take;
        cnt -= take;
    }
}

// rows [row0, row0 + nrows) of the flat key-switching key: n mask words, then b
void SerializedSwitchKey::gather_rows(u64* dst, size_t row0, size_t nrows) const {
    for (size_t r = 0; r < nrows; r++) {
        const size_t row = row0 + r;
        u64* d = dst + r * (n + 1);
        memcpy(d, base + rowA_off[row], n * 8);
        memcpy(d + n, base + rowB_off[row / dKS] + (row % dKS) * 8, 8);
    }
}

}  // namespace tfhe_b200
