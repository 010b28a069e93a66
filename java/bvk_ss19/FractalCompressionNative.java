// FractalCompressionNative.java -- the reference's static codec facade (FractalCompression.java:12-59, 230-261,
// 547-553) with the hot path delegated to libfic_b200.so.  Same names, same static configuration, same stream.
// UNCOMPILED (no JDK in the build environment); see java/README.md.
package bvk_ss19;

import java.io.DataInputStream;
import java.io.DataOutputStream;

public class FractalCompressionNative {
    public static int blockgroesse = 8;   // FractalCompression.java:14
    public static int widthKernel = 2;    // FractalCompression.java:15
    public static float avgError;         // FractalCompression.java:20: never reset between decodes
    private static float[][] imageInfo;   // FractalCompression.java:17-18

    public static float getAvgError() { return avgError; }

    public static boolean isGreyScale(RasterImage input) {  // FractalCompression.java:32-45
        for (int p : input.argb) {
            int r = (p >> 16) & 0xff, g = (p >> 8) & 0xff, b = p & 0xff;
            if (r != g || g != b) return false;
        }
        return true;
    }

    /** FractalCompression.java:54-59 -> 109-162 / 171-219: stream to `out`, returns the one-shot collage. */
    public static RasterImage encode(RasterImage input, DataOutputStream out) throws Exception {
        boolean rgb = !isGreyScale(input);
        int nr = (input.width / blockgroesse) * (input.height / blockgroesse);
        imageInfo = new float[nr][rgb ? 5 : 3];
        int[] q = new int[nr * (rgb ? 5 : 3)];
        RasterImage collage = new RasterImage(input.width, input.height);
        try {
            FicNative.encode(rgb, input.argb, input.width, input.height, blockgroesse, widthKernel, imageInfo, q);
            out.write(FicNative.stream(rgb, input.width, input.height, blockgroesse, widthKernel, q));
            out.close();  // FractalCompression.java:259
            FicNative.collage(rgb, input.argb, input.width, input.height, blockgroesse, widthKernel, imageInfo, collage.argb);
        } catch (Exception e) {
            throw e;
        } catch (Throwable t) {
            throw new Exception(t);
        }
        return collage;
    }

    /** FractalCompression.java:547-553 -> 356-421 / 430-508. */
    public static RasterImage decode(DataInputStream in) throws Exception {
        boolean rgb = in.readInt() != 0;
        int w = in.readInt(), h = in.readInt(), block = in.readInt(), wk = in.readInt();
        int n = (w / block) * (h / block) * (rgb ? 5 : 3);
        int[] q = new int[n];
        for (int i = 0; i < n; i++) q[i] = in.readInt();
        RasterImage image = new RasterImage(w, h);
        float[] avg = {avgError};
        try {
            FicNative.decode(rgb, w, h, block, wk, q, image.argb, avg);
        } catch (Exception e) {
            throw e;
        } catch (Throwable t) {
            throw new Exception(t);
        }
        avgError = avg[0];
        return image;
    }
}
