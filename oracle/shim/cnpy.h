// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for rogersce/cnpy (unpinned external
// dependency of the reference, Readme.md:32-41), used only by
// sw/data_loader.h:51-70 to read `.npz` datasets. Datasets are absent here, so loading
// aborts; the type surface exists so that the unmodified reference headers compile.
#ifndef HISPARSE_ORACLE_SHIM_CNPY_H_
#define HISPARSE_ORACLE_SHIM_CNPY_H_
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
namespace cnpy {
struct NpyArray {
    std::vector<size_t> shape;
    std::vector<char> bytes;
    template <typename T> T *data() { return reinterpret_cast<T *>(bytes.data()); }
    template <typename T> const T *data() const { return reinterpret_cast<const T *>(bytes.data()); }
};
typedef std::map<std::string, NpyArray> npz_t;
inline npz_t npz_load(std::string path) {
    std::fprintf(stderr, "cnpy shim: npz_load(%s): datasets are not available in the oracle build\n",
                 path.c_str());
    std::abort();
}
}  // namespace cnpy
#endif
