// Host-side mirror of the reference's External/array.h surface used by drivers: Array<T> with Size / Push / Resize /
// operator[] / Clear / Last, on top of std::vector.
#ifndef WF_HOST_ARRAY_H
#define WF_HOST_ARRAY_H
#include <cstddef>
#include <vector>

template <typename T> class Array {
public:
    Array() {}
    explicit Array(size_t n) : v_(n) {}
    size_t Size() const { return v_.size(); }
    size_t size() const { return v_.size(); }
    void Push(const T & x) { v_.push_back(x); }
    void Resize(size_t n) { v_.resize(n); }
    void Clear() { v_.clear(); }
    void Reserve(size_t n) { v_.reserve(n); }
    T & operator[](size_t i) { return v_[i]; }
    const T & operator[](size_t i) const { return v_[i]; }
    T & Last() { return v_.back(); }
    T * GetPtr() { return v_.data(); }
    std::vector<T> & std_vector() { return v_; }
private:
    std::vector<T> v_;
};
#endif
