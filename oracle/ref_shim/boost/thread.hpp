// TEST INFRASTRUCTURE shim: boost is absent; the reference only needs shared_ptr/make_shared.
#pragma once
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class T, class... A> std::shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
}
