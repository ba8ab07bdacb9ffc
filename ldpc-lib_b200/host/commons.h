// Host-side utilities of the drop-in driver: die(), matrix<T>, seed handling.
// Mirrors the parts of the reference's commons_portable.h / data_structures.h that the simulation
// path touches (die: commons_portable.cpp:181-189, matrix<T>: data_structures.h:10-61, seed rules:
// commons_portable.cpp:134-160) -- same names, same behaviour, new code.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

[[noreturn]] void die(char const* format, ...);
std::string format_to_string(char const* format, ...);

// settings/random_seed (main_simulation.cpp:262): 0 = take one from the random device.
extern int initial_random_seed;
void reset_random();                      // restart the noise stream (main_simulation.cpp:492)
void ensure_random_is_initialized();      // resolves seed 0, commons_portable.cpp:146-158
unsigned long long current_noise_epoch(); // bumps on every reset_random(): not used for numerics

int next_random_int(int minInclusive, int maxExclusive);   // host-side generator for the encoder's information bits

// one bit per char, the element type of the encoder's codewords (commons_portable.h:90-109)
struct bit {
    char value;
    bit() : value(0) {}
    bit(bool that) : value(that ? 1 : 0) {}
    bit(int that) : value(that ? 1 : 0) {}
    bit& operator=(bool that) { value = that ? 1 : 0; return *this; }
    bit& operator=(int that) { value = that ? 1 : 0; return *this; }
    bit& operator^=(bit that) { value ^= that.value; return *this; }
    operator bool() const { return value != 0; }
};

template <typename T>
struct matrix {
private:
    std::vector<T> contents;
    int rows, cols;
public:
    matrix() : contents(), rows(0), cols(0) {}
    matrix(int rows_, int cols_, T const& init = T()) : contents((size_t)rows_ * cols_, init), rows(rows_), cols(cols_) {}
    int n_rows() const { return rows; }
    int n_cols() const { return cols; }
    T& operator()(int row, int col)
    {
        if (row >= rows || col >= cols || row < 0 || col < 0) die("matrix out of bounds: requested (%d, %d), rows = %d, cols = %d", row, col, rows, cols);
        return contents[(size_t)row * cols + col];
    }
    T const& operator()(int row, int col) const
    {
        if (row >= rows || col >= cols || row < 0 || col < 0) die("matrix out of bounds: requested (%d, %d), rows = %d, cols = %d", row, col, rows, cols);
        return contents[(size_t)row * cols + col];
    }
    bool operator==(matrix<T> const& that) const { return rows == that.rows && cols == that.cols && contents == that.contents; }
    bool operator!=(matrix<T> const& that) const { return !(*this == that); }
};
