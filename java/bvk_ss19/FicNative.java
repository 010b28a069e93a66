// FicNative.java -- Panama FFM binding of libfic_b200.so (include/fic_b200.h).  UNCOMPILED: no JDK exists in the
// build environment; see java/README.md.  Every downcall mirrors one C entry point one to one.
package bvk_ss19;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

final class FicNative {
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.mapLibraryName("fic_b200"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    private static final MethodHandle CREATE = h("fic_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle LAST_ERROR = h("fic_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle ENCODE_GREY = h("fic_encode_grey", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle ENCODE_RGB = h("fic_encode_rgb", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle DECODE = h("fic_decode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT,
            JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle COLLAGE = h("fic_collage", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS,
            JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle STREAM_SIZE = h("fic_stream_size", FunctionDescriptor.of(JAVA_LONG, JAVA_INT, JAVA_INT,
            JAVA_INT, JAVA_INT));
    private static final MethodHandle STREAM_WRITE = h("fic_stream_write", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_INT,
            JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));

    // Multi-GPU: one context over several GPUs of the node (one upload, NCCL broadcast over NVLink, range rows sharded).
    // `-Dfic.devices=0,1,2,3` selects them; without the property the shim stays on the single-device entries above.
    private static final MethodHandle CREATE_MULTI = h("fic_create_multi", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle MULTI_LAST_ERROR = h("fic_multi_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle MULTI_ENCODE_GREY = h("fic_multi_encode_grey", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle MULTI_ENCODE_RGB = h("fic_multi_encode_rgb", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static MemorySegment multi;

    private static synchronized MemorySegment multi() throws Throwable {
        String prop = System.getProperty("fic.devices");
        if (prop == null) return null;
        if (multi == null) {
            int[] devs = java.util.Arrays.stream(prop.split(",")).mapToInt(x -> Integer.parseInt(x.trim())).toArray();
            MemorySegment out = Arena.global().allocate(ADDRESS);
            int rc = (int) CREATE_MULTI.invokeExact(Arena.global().allocateFrom(JAVA_INT, devs), devs.length, out);
            if (rc != 0) {
                MemorySegment msg = (MemorySegment) MULTI_LAST_ERROR.invokeExact(MemorySegment.NULL);
                throw new Exception("libfic_b200 error " + rc + ": " + msg.reinterpret(512).getString(0));
            }
            multi = out.get(ADDRESS, 0);
        }
        return multi;
    }

    private static MemorySegment handle;  // one context, like the reference's static state (FractalCompression.java:14-20)

    private static synchronized MemorySegment handle() throws Throwable {
        if (handle == null) {
            MemorySegment out = Arena.global().allocate(ADDRESS);
            check((int) CREATE.invokeExact(0, out), MemorySegment.NULL);
            handle = out.get(ADDRESS, 0);
        }
        return handle;
    }

    private static void check(int rc, MemorySegment h) throws Throwable {
        if (rc != 0) {  // the reference signals every failure as `throws Exception`
            MemorySegment msg = (MemorySegment) LAST_ERROR.invokeExact(h);
            throw new Exception("libfic_b200 error " + rc + ": " + msg.reinterpret(512).getString(0));
        }
    }

    /** Fills info[NR][3|5] (imageInfo / imageInfoRGB) and q[NR*3|5] (the ints writeData emits). */
    static void encode(boolean rgb, int[] argb, int w, int h, int block, int wk, float[][] info, int[] q) throws Throwable {
        int stride = rgb ? 5 : 3, nr = info.length;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = a.allocateFrom(JAVA_INT, argb);
            MemorySegment outInfo = a.allocate(JAVA_FLOAT, (long) nr * stride);
            MemorySegment outQ = a.allocate(JAVA_INT, (long) nr * stride);
            MemorySegment m = multi();
            if (m != null) {  // the same outputs, range rows sharded over the GPUs of -Dfic.devices
                int rc = (int) (rgb ? MULTI_ENCODE_RGB : MULTI_ENCODE_GREY).invokeExact(m, in, w, h, block, wk, outInfo, outQ);
                if (rc != 0) {
                    MemorySegment msg = (MemorySegment) MULTI_LAST_ERROR.invokeExact(m);
                    throw new Exception("libfic_b200 error " + rc + ": " + msg.reinterpret(512).getString(0));
                }
            } else {
                MethodHandle f = rgb ? ENCODE_RGB : ENCODE_GREY;
                check((int) f.invokeExact(handle(), in, w, h, block, wk, 0L, (long) nr, outInfo, outQ), handle());
            }
            for (int j = 0; j < nr; j++)
                for (int k = 0; k < stride; k++) info[j][k] = outInfo.getAtIndex(JAVA_FLOAT, (long) j * stride + k);
            MemorySegment.copy(outQ, JAVA_INT, 0, q, 0, nr * stride);
        }
    }

    /** Header + codes exactly as DataOutputStream.writeInt would emit them (FractalCompression.java:230-261). */
    static byte[] stream(boolean rgb, int w, int h, int block, int wk, int[] q) throws Throwable {
        long n = (long) STREAM_SIZE.invokeExact(rgb ? 1 : 0, w, h, block);
        try (Arena a = Arena.ofConfined()) {
            MemorySegment codes = a.allocateFrom(JAVA_INT, q);
            MemorySegment out = a.allocate(n);
            // host-only helper: it never touches a handle (and sets no handle error string), so a failure is reported
            // with a fixed message instead of fic_last_error -- no GPU context is created just to report it
            int rc = (int) STREAM_WRITE.invokeExact(rgb ? 1 : 0, w, h, block, wk, codes, out, n);
            if (rc != 0) throw new Exception("libfic_b200 error " + rc + ": fic_stream_write rejected its arguments");
            return out.toArray(JAVA_BYTE);
        }
    }

    /** Decoder sweeps (FractalCompression.java:378-418 / 455-505); avgError[0] is read and written. */
    static int decode(boolean rgb, int w, int h, int block, int wk, int[] q, int[] argbOut, float[] avgError) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment codes = a.allocateFrom(JAVA_INT, q);
            MemorySegment out = a.allocate(JAVA_INT, (long) w * h);
            MemorySegment avg = a.allocateFrom(JAVA_FLOAT, avgError[0]);
            MemorySegment iters = a.allocate(JAVA_INT);
            check((int) DECODE.invokeExact(handle(), rgb ? 1 : 0, w, h, block, wk, codes, 50, out, avg, iters), handle());
            MemorySegment.copy(out, JAVA_INT, 0, argbOut, 0, w * h);
            avgError[0] = avg.get(JAVA_FLOAT, 0);
            return iters.get(JAVA_INT, 0);
        }
    }

    /** getBestGeneratedCollage[RGB] (FractalCompression.java:269-347); info is rewritten in place like the reference does. */
    static void collage(boolean rgb, int[] argb, int w, int h, int block, int wk, float[][] info, int[] argbOut) throws Throwable {
        int stride = rgb ? 5 : 3, nr = info.length;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = a.allocateFrom(JAVA_INT, argb);
            MemorySegment inf = a.allocate(JAVA_FLOAT, (long) nr * stride);
            for (int j = 0; j < nr; j++)
                for (int k = 0; k < stride; k++) inf.setAtIndex(JAVA_FLOAT, (long) j * stride + k, info[j][k]);
            MemorySegment out = a.allocate(JAVA_INT, (long) w * h);
            check((int) COLLAGE.invokeExact(handle(), rgb ? 1 : 0, in, w, h, block, wk, inf, out), handle());
            MemorySegment.copy(out, JAVA_INT, 0, argbOut, 0, w * h);
            for (int j = 0; j < nr; j++) info[j][0] = inf.getAtIndex(JAVA_FLOAT, (long) j * stride);
        }
    }

    private FicNative() {}
}
