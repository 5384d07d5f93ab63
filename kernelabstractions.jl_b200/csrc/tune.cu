#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"
#include "tune.h"

namespace b200 {
static std::mutex g_mu;
static std::map<std::string, int> &table() {
    static std::map<std::string, int> t;
    return t;
}
int tune_get(const char *key, int dflt) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = table().find(key);
    return it == table().end() ? dflt : it->second;
}
void tune_set(const char *key, int value) {
    std::lock_guard<std::mutex> lk(g_mu);
    table()[key] = value;
}
}  // namespace b200

extern "C" b200_status b200_tune_set(const char *key, int32_t value) {
    if (!key) return b200::set_error(B200_ERR_INVALID_VALUE, "key is NULL");
    b200::tune_set(key, value);
    return B200_OK;
}
extern "C" b200_status b200_tune_get(const char *key, int32_t dflt, int32_t *value) {
    if (!key || !value) return b200::set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    *value = b200::tune_get(key, dflt);
    return B200_OK;
}
