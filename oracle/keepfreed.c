/* Test infrastructure (never part of the product path).
 *
 * LD_PRELOAD shim used only when tests run the fork's unmodified `alacconvert` DECODE leg
 * (oracle/_ref/alacconvert_ref): /root/reference/convert-utility/main.cu:658 free()s the magic
 * cookie before ALACDecoder::Init parses it at main.cu:708 (use-after-free), so
 * glibc's free-list links overwrite frameLength / bitDepth / pb / mb / kb.  This shim turns
 * free() of SMALL blocks (<= 64 usable bytes: the 24- or 48-byte cookie) into a no-op, so the
 * same binary reads the cookie it just read from the file.  Larger blocks are released as usual. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <malloc.h>
#include <stddef.h>

void free(void *p)
{
    static void (*real_free)(void *) = 0;
    if (!p) return;
    if (!real_free) real_free = (void (*)(void *))dlsym(RTLD_NEXT, "free");
    if (malloc_usable_size(p) <= 64) return; /* keep the cookie (and other tiny blocks) alive */
    real_free(p);
}
