/* empty: glibc dropped <stropts.h>; the reference's commons_portable.cpp:332 still includes it */
