#include "commons.h"

#include <random>

int initial_random_seed = 0;
static bool random_initialized = false;
static unsigned long long noise_epoch = 0;

void die(char const* format, ...)
{
    va_list ap;
    va_start(ap, format);
    vfprintf(stderr, format, ap);
    va_end(ap);
    fputs("\n", stderr);
    exit(1);
}

std::string format_to_string(char const* format, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, format);
    vsnprintf(buf, sizeof buf, format, ap);
    va_end(ap);
    return buf;
}

void reset_random()
{
    random_initialized = false;
    ++noise_epoch;
}

void ensure_random_is_initialized()
{
    if (random_initialized) return;
    if (initial_random_seed == 0) {
        std::random_device device;
        initial_random_seed = (int)device();
        if (initial_random_seed == 0) initial_random_seed = 1;
    }
    random_initialized = true;
}

unsigned long long current_noise_epoch() { return noise_epoch; }

// The encoder's information bits come from a host generator seeded like the reference's (commons_portable.cpp:134-166);
// the GPU noise does not consume it.
static std::mt19937 host_generator(1);
static bool host_generator_seeded = false;
int next_random_int(int minInclusive, int maxExclusive)
{
    ensure_random_is_initialized();
    if (!host_generator_seeded) { host_generator = std::mt19937(initial_random_seed); host_generator_seeded = true; }
    std::uniform_int_distribution<int> dist(minInclusive, maxExclusive - 1);
    return dist(host_generator);
}
