// helpers/insert.hpp -- insert-or-overwrite into the model's coefficient maps
// (reference: /root/reference/include/helpers/insert.hpp:27-38; pinned by
// tests/qubo_test.cpp:25-27 "add_variable overwrites").
#ifndef ONESOLVER_B200_HELPERS_INSERT_HPP_
#define ONESOLVER_B200_HELPERS_INSERT_HPP_

#include <unordered_map>

namespace helpers {

template <class Key, class Coef, class Hash>
void insert_model(std::unordered_map<Key, Coef, Hash> &store, const Key &key, const Coef &value) {
  store.insert_or_assign(key, value);
}

}  // namespace helpers

#endif
