// placeholder until the 2D path lands
#pragma once
#include "gdk_ctx.h"
extern "C" int32_t gdk_density2d_batch(gdk_ctx* ctx, int32_t, const gdk_spec2d*, double*, const int64_t*, gdk_result2d*, uint32_t) {
    return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "2D path not built yet");
}
extern "C" int32_t gdk_hist2d_batch(gdk_ctx* ctx, int32_t, const gdk_spec2d*, double*, const int64_t*) {
    return gdk_fail(ctx, GDK_ERR_UNSUPPORTED, "2D path not built yet");
}
