// HeadlessMain.java -- what RLEAppController.openDecodedImage (RLEAppController.java:172-188) does, without JavaFX:
// encode the image to unknown.run, decode the stream, print the label the GUI shows.  UNCOMPILED; see java/README.md.
package bvk_ss19;

import java.awt.image.BufferedImage;
import java.io.DataInputStream;
import java.io.DataOutputStream;
import java.io.File;
import java.io.FileInputStream;
import java.io.FileOutputStream;
import javax.imageio.ImageIO;

public class HeadlessMain {
    public static void main(String[] args) throws Exception {
        if (args.length < 1) {
            System.err.println("usage: HeadlessMain image [blockgroesse] [widthKernel]");
            System.exit(2);
        }
        if (args.length > 1) FractalCompressionNative.blockgroesse = Integer.parseInt(args[1]);
        if (args.length > 2) FractalCompressionNative.widthKernel = Integer.parseInt(args[2]);
        BufferedImage img = ImageIO.read(new File(args[0]));
        RasterImage source = new RasterImage(img.getWidth(), img.getHeight());
        img.getRGB(0, 0, source.width, source.height, source.argb, 0, source.width);  // TYPE_INT_ARGB, like RasterImage.java:44
        FractalCompressionNative.encode(source, new DataOutputStream(new FileOutputStream("unknown.run")));
        RasterImage decoded = FractalCompressionNative.decode(new DataInputStream(new FileInputStream("unknown.run")));
        BufferedImage out = new BufferedImage(decoded.width, decoded.height, BufferedImage.TYPE_INT_ARGB);
        out.setRGB(0, 0, decoded.width, decoded.height, decoded.argb, 0, decoded.width);
        ImageIO.write(out, "png", new File("decoded.png"));
        System.out.println("MSE " + FractalCompressionNative.getAvgError());  // RLEAppController.java:180
    }
}
