// Standalone ASan/UBSan driver for the host lanes (developer check; not part of the product).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "/root/repo/include/dyd.h"
static std::vector<uint8_t> slurp(const char* p) { std::ifstream f(p, std::ios::binary); return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), {}); }
int main(int argc, char** argv) {
    // argv[1]: cells file: u64 n, then n x (u64 len, bytes); argv[2]: csv file
    auto raw = slurp(argv[1]);
    const uint8_t* p = raw.data();
    uint64_t n; memcpy(&n, p, 8); p += 8;
    std::vector<int64_t> off(n + 1, 0); std::vector<uint8_t> text;
    for (uint64_t i = 0; i < n; ++i) { uint64_t l; memcpy(&l, p, 8); p += 8; text.insert(text.end(), p, p + l); p += l; off[i + 1] = (int64_t)text.size(); }
    text.push_back(0);
    long checks = 0;
    for (int mode = 0; mode <= 2; ++mode) {
        dyd_ingest* h = nullptr;
        if (dyd_ingest_cells(text.data(), off.data(), nullptr, (int64_t)n, mode, 4, &h)) { printf("ingest failed\n"); return 1; }
        int64_t no, nv, ns; dyd_ingest_sizes(h, &no, &nv, &ns);
        int64_t nc = 0; dyd_ingest_effective_text(h, nullptr, nullptr, &nc, nullptr, nullptr, 0);
        std::vector<int64_t> eoff(off); std::vector<uint8_t> etext(text);
        if (nc) { eoff.assign(n + 1, 0); dyd_ingest_effective_text(h, text.data(), off.data(), nullptr, eoff.data(), nullptr, 4); etext.assign((size_t)eoff[n] + 1, 0); dyd_ingest_effective_text(h, text.data(), off.data(), nullptr, eoff.data(), etext.data(), 4); }
        std::vector<uint8_t> status(n);
        if (mode == 0) {
            std::vector<int64_t> img(n + 1), poly(no + 1), who(2 * n); std::vector<double> xy(2 * nv + 2); std::vector<int32_t> whl(2 * n); std::vector<uint8_t> whk(2 * n);
            dyd_ingest_export_polygons(h, status.data(), img.data(), poly.data(), xy.data(), who.data(), whl.data(), whk.data(), 4);
            std::vector<int32_t> arg(4 * no + 4, 0); std::vector<uint8_t> valid(no + 1, 0);
            for (int64_t q = 0; q < no; ++q) { const int64_t V = poly[q + 1] - poly[q]; valid[q] = V > 0; if (V > 0) { arg[4 * q] = 0; arg[4 * q + 1] = (int32_t)(V - 1); arg[4 * q + 2] = (int32_t)(V / 2); arg[4 * q + 3] = 0; } }
            std::vector<int64_t> oo(n + 1);
            dyd_egress_ptlist(h, etext.data(), eoff.data(), arg.data(), valid.data(), oo.data(), nullptr, 4);
            std::vector<uint8_t> out((size_t)oo[n] + 1);
            dyd_egress_ptlist(h, etext.data(), eoff.data(), arg.data(), valid.data(), oo.data(), out.data(), 4);
            checks += oo[n];
        } else if (mode == 1) {
            std::vector<int64_t> img(n + 1); std::vector<double> pts(4 * no + 4); std::vector<uint8_t> valid(no + 1);
            dyd_ingest_export_boxes(h, status.data(), img.data(), pts.data(), valid.data(), 4);
            checks += no;
        } else {
            std::vector<int64_t> cell(n + 1), noff(no + 1), ooff(no + 1); std::vector<int32_t> nlen(no + 1), olen(no + 1), ll(n);
            dyd_ingest_export_names(h, status.data(), cell.data(), noff.data(), nlen.data(), 4);
            dyd_ingest_export_objects(h, ll.data(), ooff.data(), olen.data(), 4);
            const char* vocab = "newname\"x\\";    // two entries
            std::vector<uint8_t> vb(vocab, vocab + 10); std::vector<int64_t> vo = {0, 7, 10};
            std::vector<uint8_t> flag(no + 1, 0); std::vector<int32_t> nid(no + 1, 0);
            for (int64_t q = 0; q < no; ++q) { flag[q] = q % 3 != 0; nid[q] = (int32_t)(q & 1); }
            std::vector<int64_t> oo(n + 1);
            dyd_egress_names(h, etext.data(), eoff.data(), flag.data(), nid.data(), vb.data(), vo.data(), 2, oo.data(), nullptr, 4);
            std::vector<uint8_t> out((size_t)oo[n] + 1);
            dyd_egress_names(h, etext.data(), eoff.data(), flag.data(), nid.data(), vb.data(), vo.data(), 2, oo.data(), out.data(), 4);
            // expanded rows: every named object once
            std::vector<int64_t> ec, eo; std::vector<int32_t> et;
            for (uint64_t r = 0; r < n; ++r) for (int64_t q = cell[r]; q < cell[r + 1]; ++q) if (nlen[q] >= 0) { ec.push_back((int64_t)r); eo.push_back(q); et.push_back((int32_t)(q & 1)); }
            std::vector<int64_t> so(ec.size() + 1);
            dyd_egress_split(h, etext.data(), eoff.data(), (int64_t)ec.size(), ec.data(), eo.data(), et.data(), vb.data(), vo.data(), 2, so.data(), nullptr, 4);
            std::vector<uint8_t> sout((size_t)so[ec.size()] + 1);
            dyd_egress_split(h, etext.data(), eoff.data(), (int64_t)ec.size(), ec.data(), eo.data(), et.data(), vb.data(), vo.data(), 2, so.data(), sout.data(), 4);
            checks += oo[n] + so[ec.size()];
        }
        dyd_ingest_free(h);
    }
    // canonical rewriter on every cell
    for (uint64_t i = 0; i < n; ++i) { std::vector<uint8_t> o((size_t)(off[i + 1] - off[i]) * 6 + 64); checks += dyd_json_canonical(text.data() + off[i], off[i + 1] - off[i], o.data(), (int64_t)o.size()) > 0; }
    // CSV reader
    auto csv = slurp(argv[2]);
    const char* nas = "NAnull"; std::vector<int64_t> na_off = {0, 0, 2, 6};
    void* hc = nullptr;
    dyd_csv_open(csv.data(), (int64_t)csv.size(), (const uint8_t*)nas, na_off.data(), 3, 4, &hc);
    int64_t nr; int32_t ncol, fl; int64_t hb, he; dyd_csv_info(hc, &nr, &ncol, &hb, &he, &fl);
    if (!(fl & 1) && nr > 0) {
        std::vector<int64_t> cb(ncol), cn(ncol); std::vector<uint8_t> ct(ncol), cu(ncol);
        dyd_csv_measure(hc, 7, cb.data(), cn.data(), ct.data(), cu.data(), 4);
        std::vector<std::vector<int64_t>> offs(ncol, std::vector<int64_t>(nr + 1)); std::vector<std::vector<uint8_t>> datas(ncol), maps(ncol);
        std::vector<int32_t> cols(ncol); std::vector<int64_t*> po(ncol); std::vector<uint8_t*> pd(ncol), pm(ncol);
        for (int j = 0; j < ncol; ++j) { cols[j] = j; datas[j].resize((size_t)cb[j] + 1); maps[j].resize((size_t)(nr + 7) / 8); po[j] = offs[j].data(); pd[j] = datas[j].data(); pm[j] = maps[j].data(); }
        dyd_csv_fill(hc, ncol, cols.data(), po.data(), pd.data(), pm.data(), 4);
        checks += nr * ncol;
        // round-trip verdict of every column, then the straight-to-file writer (all rows, a row selection) and the wide tokenizer
        // on what it wrote: the cells must come back unchanged
        for (int j = 0; j < ncol; ++j) checks += dyd_csv_roundtrip_check(po[j], pd[j], nullptr, nr, 5, (const uint8_t*)nas, na_off.data(), 3, 1, 4) >= 0;
        if (ncol >= 2) {
            std::vector<int32_t> kinds(ncol, 0);
            std::vector<const int64_t*> co(ncol); std::vector<const uint8_t*> cd(ncol), cv(ncol, nullptr);
            std::vector<std::vector<uint8_t>> vbytes(ncol);
            for (int j = 0; j < ncol; ++j) {
                co[j] = po[j]; cd[j] = pd[j];
                vbytes[j].resize((size_t)nr);
                for (int64_t r = 0; r < nr; ++r) vbytes[j][(size_t)r] = (maps[j][(size_t)r / 8] >> (r % 8)) & 1;
                cv[j] = vbytes[j].data();
            }
            std::string hdr = "\xef\xbb\xbf";
            for (int j = 0; j < ncol; ++j) { hdr += (j ? ",c" : "c") + std::to_string(j); }
            hdr += "\n";
            std::vector<int64_t> rows;
            for (int64_t r = nr - 1; r >= 0; r -= 3) rows.push_back(r);
            int64_t wrote = 0;
            const char* out_path = "/tmp/dyd_asan_out.csv";
            if (dyd_csv_write_file(out_path, 0, (const uint8_t*)hdr.data(), (int64_t)hdr.size(), kinds.data(), co.data(), cd.data(), cv.data(), ncol, rows.data(),
                                   (int64_t)rows.size(), 3, &wrote)) { printf("write (rows) failed\n"); return 1; }
            if (dyd_csv_write_file(out_path, 0, (const uint8_t*)hdr.data(), (int64_t)hdr.size(), kinds.data(), co.data(), cd.data(), cv.data(), ncol, nullptr, nr, 3,
                                   &wrote)) { printf("write failed\n"); return 1; }
            std::vector<uint8_t> back((size_t)wrote);
            if (dyd_read_file(out_path, back.data(), wrote, 3)) { printf("read back failed\n"); return 1; }
            void* h2 = nullptr;
            dyd_csv_open(back.data(), wrote, (const uint8_t*)nas, na_off.data(), 3, 3, &h2);
            int64_t nr2; int32_t nc2, fl2; dyd_csv_info(h2, &nr2, &nc2, &hb, &he, &fl2);
            if (!(fl2 & 1) && nr2 > 0) {
                std::vector<int64_t> cb2(nc2), cn2(nc2); std::vector<uint8_t> ct2(nc2), cu2(nc2);
                dyd_csv_measure(h2, 7, cb2.data(), cn2.data(), ct2.data(), cu2.data(), 3);
                std::vector<std::vector<int64_t>> o2(nc2, std::vector<int64_t>(nr2 + 1)); std::vector<std::vector<uint8_t>> d2(nc2), m2(nc2);
                std::vector<int32_t> c2(nc2); std::vector<int64_t*> p2(nc2); std::vector<uint8_t*> q2(nc2), r2(nc2);
                for (int j = 0; j < nc2; ++j) { c2[j] = j; d2[j].resize((size_t)std::max<int64_t>(cb2[j], 1)); m2[j].resize((size_t)(nr2 + 7) / 8); p2[j] = o2[j].data(); q2[j] = d2[j].data(); r2[j] = m2[j].data(); }
                dyd_csv_fill(h2, nc2, c2.data(), p2.data(), q2.data(), r2.data(), 3);
                // rows whose every cell is missing are blank lines in the file and vanish on the way back: compare only when none did
                if (nr2 == nr && nc2 == ncol)
                    for (int j = 0; j < ncol; ++j)
                        if (cb2[j] != cb[j] || memcmp(d2[j].data(), datas[j].data(), (size_t)cb[j]) != 0) { printf("round trip changed column %d (flags %d)\n", j, fl2); return 1; }
                checks += nr2 + (fl2 & 4);
            }
            dyd_csv_close(h2);
        }
    }
    dyd_csv_close(hc);
    printf("ok %ld\n", checks);
    return 0;
}
