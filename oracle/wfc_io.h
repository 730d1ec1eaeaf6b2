// TEST INFRASTRUCTURE -- shared by the oracle programs only (oracle/ref_driver.cpp, oracle/wf_oracle.cpp).
// "WFC1" record container used for oracle case files and state dumps. Python twin: tests/wfc.py.
//   file   := magic "WFC1" , record*
//   record := char name[32] (NUL padded) , int32 dtype (0 = float64, 1 = int32, 2 = int64) , int64 count , data
#ifndef WF_ORACLE_WFC_IO_H
#define WF_ORACLE_WFC_IO_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace wfc {

struct Record {
    int dtype;
    std::vector<double> f64;
    std::vector<int32_t> i32;
    std::vector<int64_t> i64;
    size_t size() const { return dtype == 0 ? f64.size() : (dtype == 1 ? i32.size() : i64.size()); }
};

struct File {
    std::map<std::string, Record> rec;
    std::vector<std::string> order;

    bool has(const std::string & k) const { return rec.count(k) != 0; }
    const std::vector<double> & f(const std::string & k) const {
        std::map<std::string, Record>::const_iterator it = rec.find(k);
        if (it == rec.end() || it->second.dtype != 0) { std::fprintf(stderr, "wfc: missing f64 record '%s'\n", k.c_str()); std::exit(3); }
        return it->second.f64;
    }
    const std::vector<int32_t> & i(const std::string & k) const {
        std::map<std::string, Record>::const_iterator it = rec.find(k);
        if (it == rec.end() || it->second.dtype != 1) { std::fprintf(stderr, "wfc: missing i32 record '%s'\n", k.c_str()); std::exit(3); }
        return it->second.i32;
    }
    double scalar(const std::string & k, double dflt) const {
        std::map<std::string, Record>::const_iterator it = rec.find(k);
        if (it == rec.end()) return dflt;
        if (it->second.dtype == 0) return it->second.f64.at(0);
        if (it->second.dtype == 1) return (double)it->second.i32.at(0);
        return (double)it->second.i64.at(0);
    }
    void put_f(const std::string & k, const double * p, size_t n) { Record r; r.dtype = 0; r.f64.assign(p, p + n); add(k, r); }
    void put_i(const std::string & k, const int32_t * p, size_t n) { Record r; r.dtype = 1; r.i32.assign(p, p + n); add(k, r); }
    void put_l(const std::string & k, const int64_t * p, size_t n) { Record r; r.dtype = 2; r.i64.assign(p, p + n); add(k, r); }
    void put_scalar(const std::string & k, double v) { put_f(k, &v, 1); }
    void add(const std::string & k, const Record & r) { if (!rec.count(k)) order.push_back(k); rec[k] = r; }

    bool read(const char * path) {
        FILE * fp = std::fopen(path, "rb"); if (!fp) return false;
        char magic[4]; if (std::fread(magic, 1, 4, fp) != 4 || std::memcmp(magic, "WFC1", 4)) { std::fclose(fp); return false; }
        for (;;) {
            char name[32]; if (std::fread(name, 1, 32, fp) != 32) break;
            int32_t dt; int64_t cnt;
            if (std::fread(&dt, 4, 1, fp) != 1 || std::fread(&cnt, 8, 1, fp) != 1) { std::fclose(fp); return false; }
            Record r; r.dtype = dt; size_t got = 0;
            if (dt == 0) { r.f64.resize(cnt); got = std::fread(r.f64.data(), 8, cnt, fp); }
            else if (dt == 1) { r.i32.resize(cnt); got = std::fread(r.i32.data(), 4, cnt, fp); }
            else { r.i64.resize(cnt); got = std::fread(r.i64.data(), 8, cnt, fp); }
            if ((int64_t)got != cnt) { std::fclose(fp); return false; }
            name[31] = 0; add(std::string(name), r);
        }
        std::fclose(fp); return true;
    }
    bool write(const char * path) const {
        FILE * fp = std::fopen(path, "wb"); if (!fp) return false;
        std::fwrite("WFC1", 1, 4, fp);
        for (size_t k = 0; k < order.size(); ++k) {
            const Record & r = rec.find(order[k])->second;
            char name[32]; std::memset(name, 0, 32); std::strncpy(name, order[k].c_str(), 31);
            int32_t dt = r.dtype; int64_t cnt = (int64_t)r.size();
            std::fwrite(name, 1, 32, fp); std::fwrite(&dt, 4, 1, fp); std::fwrite(&cnt, 8, 1, fp);
            if (dt == 0) std::fwrite(r.f64.data(), 8, cnt, fp);
            else if (dt == 1) std::fwrite(r.i32.data(), 4, cnt, fp);
            else std::fwrite(r.i64.data(), 8, cnt, fp);
        }
        std::fclose(fp); return true;
    }
};

} // namespace wfc
#endif
