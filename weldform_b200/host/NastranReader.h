// Host-side mirror of SPH::NastranReader (Source/NastranReader.h): reads GRID / CTRIA3 cards of a small-field (8-column)
// Nastran bulk deck into node coordinates and triangle connectivity, which is all TriMesh(NastranReader&) consumes
// (Source/Mesh.cpp:17-60). Host I/O only.
#ifndef WF_HOST_NASTRAN_READER_H
#define WF_HOST_NASTRAN_READER_H
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>
#include "fatal.h"

namespace SPH {

class NastranReader {
public:
    int dim = 3, node_count = 0, elem_count = 0;
    std::vector<double> node;      // 3 * node_count
    std::vector<int> elcon;        // 3 * elem_count, node POSITIONS (0-based)
    NastranReader() {}
    NastranReader(const char * fName) { read(fName); }
    static double field_to_double(std::string f)
    {
        // Nastran reals may omit the exponent letter: 1.5-3 == 1.5e-3
        size_t b = f.find_first_not_of(' '); if (b == std::string::npos) return 0.0;
        f = f.substr(b); size_t e = f.find_last_not_of(' '); f = f.substr(0, e + 1);
        for (size_t i = 1; i < f.size(); ++i)
            if ((f[i] == '-' || f[i] == '+') && f[i - 1] != 'e' && f[i - 1] != 'E' && f[i - 1] != 'd' && f[i - 1] != 'D') { f.insert(i, "e"); break; }
        for (auto & ch : f) if (ch == 'd' || ch == 'D') ch = 'e';
        return std::atof(f.c_str());
    }
    void read(const char * fName)
    {
        std::ifstream file(fName);
        if (!file) throw new Fatal("NastranReader: cannot open %s", fName);
        std::string line; std::map<int, int> nodepos; std::vector<int> raw;
        while (std::getline(file, line)) {
            if (line.size() < 8) continue;
            line.resize(80, ' ');
            const std::string card = line.substr(0, 8);
            if (card.compare(0, 4, "GRID") == 0 && card.compare(0, 5, "GRID*") != 0) {
                const int id = std::atoi(line.substr(8, 8).c_str());
                nodepos[id] = node_count++;
                for (int k = 0; k < 3; ++k) node.push_back(field_to_double(line.substr(24 + 8 * k, 8)));
            } else if (card.compare(0, 6, "CTRIA3") == 0) {
                for (int k = 0; k < 3; ++k) raw.push_back(std::atoi(line.substr(24 + 8 * k, 8).c_str()));
                ++elem_count;
            }
        }
        elcon.resize(raw.size());
        for (size_t i = 0; i < raw.size(); ++i) {
            auto it = nodepos.find(raw[i]);
            if (it == nodepos.end()) throw new Fatal("NastranReader: element refers to unknown node %d", raw[i]);
            elcon[i] = it->second;
        }
    }
};

} // namespace SPH
#endif
