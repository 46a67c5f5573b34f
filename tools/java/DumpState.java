// DumpState.java -- writes jxlatte's post-entropy frame state and its reconstructed planes, so that oracle/ (and through it the
// CUDA path) can be pinned against a real JVM run of the reference.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no JDK here or on the GPU boxes (DESIGN.md, section 2).  It is written against
// the reference tree as it is (java/com/traneptora/jxlatte, field names as of the commit under /root/reference) and uses only
// java.lang.reflect, java.io and java.nio.  tests/golden/from_jvm/README.md has the three one-line hooks to add to the reference
// and the two commands to run; tests/test_from_jvm.py consumes whatever this writes.
//
// Drop this file into java/com/traneptora/jxlatte/util/ of the reference and add the hooks:
//   Frame.decodePassGroups(), after the invertVarDCT loop (Frame.java:374):         DumpState.afterPasses(this, passGroups);
//   Frame.decodeFrame(), after performEdgePreservingFilter (Frame.java:461):        DumpState.afterFilters(this);
//   JXLCodestreamDecoder.decode(), after performColorTransforms(matrix, frame) (JXLCodestreamDecoder.java:637):
//                                                                                   DumpState.afterColor(frame, matrix, imageHeader);
// Nothing is written unless the JVM runs with -Djxlatte.dump=<directory>.
//
// Output (little endian, row-major, one file per array, frame index k in the name; manifest.json lists shape and dtype):
//   f<k>_qcoeff_<c>.i32     (H>>sy) x (W>>sx)  HFCoefficients.quantizedCoeffs of the LAST pass (passes already summed by invertVarDCT),
//                           stitched over groups; c = 0,1,2 in frame order X,Y,B
//   f<k>_lf_<c>.f32         dequantLFCoeff stitched over LF groups
//   f<k>_dct_select.u8, f<k>_block_origin.u8, f<k>_hf_mul.i32, f<k>_sharpness.i32, f<k>_x_from_y.i32, f<k>_b_from_y.i32
//   f<k>_qm_weights.f32     HFGlobal.weights flattened [17][3][matrixH][matrixW] (the layout of jxlb200_qm_generate)
//   f<k>_after_idct_<c>.f32, f<k>_after_filters_<c>.f32, f<k>_after_color_<c>.f32   the reference's own planes (padded size)
//   f<k>_params.json        the scalars of jxlb200_frame_params
package com.traneptora.jxlatte.util;

import java.io.FileOutputStream;
import java.io.IOException;
import java.io.PrintWriter;
import java.lang.reflect.Field;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.channels.FileChannel;

import com.traneptora.jxlatte.bundle.ImageHeader;
import com.traneptora.jxlatte.color.OpsinInverseMatrix;
import com.traneptora.jxlatte.frame.Frame;
import com.traneptora.jxlatte.frame.FrameFlags;
import com.traneptora.jxlatte.frame.FrameHeader;
import com.traneptora.jxlatte.frame.LFGlobal;
import com.traneptora.jxlatte.frame.features.RestorationFilter;
import com.traneptora.jxlatte.frame.group.LFGroup;
import com.traneptora.jxlatte.frame.group.PassGroup;
import com.traneptora.jxlatte.frame.vardct.HFGlobal;
import com.traneptora.jxlatte.frame.vardct.TransformType;

public final class DumpState {
    private static final String DIR = System.getProperty("jxlatte.dump");
    private static int frameIndex = -1;
    private static StringBuilder manifest = new StringBuilder();

    private DumpState() {}

    private static Object get(Object o, String name) {
        try {
            Class<?> c = o.getClass();
            while (c != null) {
                try {
                    Field f = c.getDeclaredField(name);
                    f.setAccessible(true);
                    return f.get(o);
                } catch (NoSuchFieldException e) {
                    c = c.getSuperclass();
                }
            }
            throw new IllegalStateException("no field " + name);
        } catch (IllegalAccessException e) {
            throw new IllegalStateException(e);
        }
    }

    private static void write(String name, ByteBuffer bb, String dtype, int h, int w) {
        bb.flip();
        try (FileChannel ch = new FileOutputStream(DIR + "/" + name).getChannel()) {
            while (bb.hasRemaining())
                ch.write(bb);
        } catch (IOException e) {
            throw new IllegalStateException(e);
        }
        manifest.append(String.format("{\"file\": \"%s\", \"dtype\": \"%s\", \"shape\": [%d, %d]},%n", name, dtype, h, w));
    }

    private static ByteBuffer buf(long bytes) {
        return ByteBuffer.allocate((int)bytes).order(ByteOrder.LITTLE_ENDIAN);
    }

    private static void writeFloats(String name, float[][] a, int h, int w) {
        ByteBuffer bb = buf(4L * h * w);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                bb.putFloat(a[y][x]);
        write(name, bb, "f32", h, w);
    }

    private static void writeInts(String name, int[][] a, int h, int w) {
        ByteBuffer bb = buf(4L * h * w);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                bb.putInt(a[y][x]);
        write(name, bb, "i32", h, w);
    }

    private static void writeBytes(String name, byte[][] a, int h, int w) {
        ByteBuffer bb = buf((long)h * w);
        for (int y = 0; y < h; y++)
            bb.put(a[y], 0, w);
        write(name, bb, "u8", h, w);
    }

    private static void planes(Frame frame, String tag) {
        ImageBuffer[] buffer = frame.getBuffer();
        Dimension padded = frame.getPaddedFrameSize();
        for (int c = 0; c < 3; c++) {
            if (!buffer[c].isFloat())
                return;
            writeFloats(String.format("f%d_%s_%d.f32", frameIndex, tag, c), buffer[c].getFloatBuffer(),
                Math.min(padded.height, buffer[c].height), Math.min(padded.width, buffer[c].width));
        }
    }

    /** Frame.decodePassGroups, after the invertVarDCT loop: the inputs of jxlb200_vardct_reconstruct and the planes after stage 1. */
    public static void afterPasses(Frame frame, PassGroup[][] passGroups) {
        if (DIR == null)
            return;
        frameIndex++;
        FrameHeader header = frame.getFrameHeader();
        if (header.encoding != FrameFlags.VARDCT)
            return;
        Dimension padded = frame.getPaddedFrameSize();
        int H = padded.height, W = padded.width;
        LFGlobal lfGlobal = (LFGlobal)get(frame, "lfGlobal");
        HFGlobal hfGlobal = (HFGlobal)get(frame, "hfGlobal");
        LFGroup[] lfGroups = (LFGroup[])get(frame, "lfGroups");
        int numGroups = (Integer)get(frame, "numGroups");
        int last = passGroups.length - 1;

        // coefficients: the last pass holds the sum of all passes (PassGroup.invertVarDCT :174-200)
        for (int c = 0; c < 3; c++) {
            int sy = header.jpegUpsamplingY[c], sx = header.jpegUpsamplingX[c];
            int[][] q = new int[H >> sy][W >> sx];
            for (int g = 0; g < numGroups; g++) {
                if (passGroups[last][g].hfCoefficients == null)
                    continue;
                int[][] src = passGroups[last][g].hfCoefficients.quantizedCoeffs[c];
                Point loc = frame.getGroupLocation(g);                       // in groups (y, x)
                int oy = (loc.y * header.groupDim) >> sy, ox = (loc.x * header.groupDim) >> sx;
                for (int y = 0; y < src.length && oy + y < q.length; y++)
                    System.arraycopy(src[y], 0, q[oy + y], ox, Math.min(src[y].length, q[0].length - ox));
            }
            writeInts(String.format("f%d_qcoeff_%d.i32", frameIndex, c), q, H >> sy, W >> sx);
        }
        // LF planes and the per-block maps, stitched over LF groups (2048 x 2048 px = 256 x 256 blocks)
        int hb = H >> 3, wb = W >> 3, th = (H + 63) >> 6, tw = (W + 63) >> 6;
        byte[][] dct = new byte[hb][wb], org = new byte[hb][wb];
        int[][] hfm = new int[hb][wb], sharp = new int[hb][wb], xfy = new int[th][tw], bfy = new int[th][tw];
        float[][][] lf = new float[3][][];
        for (int c = 0; c < 3; c++)
            lf[c] = new float[hb >> header.jpegUpsamplingY[c]][wb >> header.jpegUpsamplingX[c]];
        for (LFGroup lfg : lfGroups) {
            Point loc = frame.getLFGroupLocation(lfg.lfGroupID);           // in LF groups (y, x)
            int by = loc.y << 8, bx = loc.x << 8;
            for (int c = 0; c < 3; c++) {
                float[][] src = lfg.lfCoeff.dequantLFCoeff[c];
                int oy = by >> header.jpegUpsamplingY[c], ox = bx >> header.jpegUpsamplingX[c];
                for (int y = 0; y < src.length && oy + y < lf[c].length; y++)
                    System.arraycopy(src[y], 0, lf[c][oy + y], ox, Math.min(src[y].length, lf[c][0].length - ox));
            }
            TransformType[][] ds = lfg.hfMetadata.dctSelect;
            for (int y = 0; y < ds.length && by + y < hb; y++)
                for (int x = 0; x < ds[y].length && bx + x < wb; x++) {
                    dct[by + y][bx + x] = (byte)(ds[y][x] == null ? 0 : ds[y][x].type);
                    hfm[by + y][bx + x] = lfg.hfMetadata.hfMultiplier[y][x];
                    sharp[by + y][bx + x] = lfg.hfMetadata.hfStreamBuffer[3][y][x];
                }
            for (Point p : lfg.hfMetadata.blockList)
                if (by + p.y < hb && bx + p.x < wb)
                    org[by + p.y][bx + p.x] = 1;
            int[][] xf = lfg.hfMetadata.hfStreamBuffer[0], bf = lfg.hfMetadata.hfStreamBuffer[1];
            int ty = loc.y << 5, tx = loc.x << 5;                           // 64 x 64 px tiles: 32 per LF group side
            for (int y = 0; y < xf.length && ty + y < th; y++)
                for (int x = 0; x < xf[y].length && tx + x < tw; x++) {
                    xfy[ty + y][tx + x] = xf[y][x];
                    bfy[ty + y][tx + x] = bf[y][x];
                }
        }
        for (int c = 0; c < 3; c++)
            writeFloats(String.format("f%d_lf_%d.f32", frameIndex, c), lf[c], lf[c].length, lf[c][0].length);
        writeBytes(String.format("f%d_dct_select.u8", frameIndex), dct, hb, wb);
        writeBytes(String.format("f%d_block_origin.u8", frameIndex), org, hb, wb);
        writeInts(String.format("f%d_hf_mul.i32", frameIndex), hfm, hb, wb);
        writeInts(String.format("f%d_sharpness.i32", frameIndex), sharp, hb, wb);
        writeInts(String.format("f%d_x_from_y.i32", frameIndex), xfy, th, tw);
        writeInts(String.format("f%d_b_from_y.i32", frameIndex), bfy, th, tw);
        // quant weights: [parameter index][channel][matrixH][matrixW], the order jxlb200_qm_generate uses
        {
            int total = 0;
            for (float[][][] p : hfGlobal.weights)
                for (float[][] ch : p)
                    total += ch.length * ch[0].length;
            ByteBuffer bb = buf(4L * total);
            for (float[][][] p : hfGlobal.weights)
                for (float[][] ch : p)
                    for (float[] row : ch)
                        for (float v : row)
                            bb.putFloat(v);
            write(String.format("f%d_qm_weights.f32", frameIndex), bb, "f32", 1, total);
        }
        planes(frame, "after_idct");
        // scalars
        RestorationFilter rf = header.restorationFilter;
        OpsinInverseMatrix oim = frame.globalMetadata.getOpsinInverseMatrix();
        try (PrintWriter pw = new PrintWriter(DIR + "/" + String.format("f%d_params.json", frameIndex))) {
            pw.printf("{\"width\": %d, \"height\": %d, \"global_scale\": %d, \"xqm_scale\": %d, \"bqm_scale\": %d,%n", W, H,
                lfGlobal.globalScale, header.xqmScale, header.bqmScale);
            pw.printf(" \"quant_bias\": [%s, %s, %s], \"quant_bias_numerator\": %s,%n", bits(oim.quantBias[0]), bits(oim.quantBias[1]),
                bits(oim.quantBias[2]), bits(oim.quantBiasNumerator));
            pw.printf(" \"color_factor\": %d, \"base_corr_x\": %s, \"base_corr_b\": %s,%n", lfGlobal.lfChanCorr.colorFactor,
                bits(lfGlobal.lfChanCorr.baseCorrelationX), bits(lfGlobal.lfChanCorr.baseCorrelationB));
            pw.printf(" \"shift_x\": [%d, %d, %d], \"shift_y\": [%d, %d, %d],%n", header.jpegUpsamplingX[0], header.jpegUpsamplingX[1],
                header.jpegUpsamplingX[2], header.jpegUpsamplingY[0], header.jpegUpsamplingY[1], header.jpegUpsamplingY[2]);
            pw.printf(" \"gab\": %d, \"gab_w1\": [%s, %s, %s], \"gab_w2\": [%s, %s, %s],%n", rf.gab ? 1 : 0, bits(rf.gab1Weights[0]),
                bits(rf.gab1Weights[1]), bits(rf.gab1Weights[2]), bits(rf.gab2Weights[0]), bits(rf.gab2Weights[1]), bits(rf.gab2Weights[2]));
            pw.printf(" \"epf_iters\": %d, \"epf_sharp_lut\": [", rf.epfIterations);
            for (int i = 0; i < 8; i++)
                pw.printf("%s%s", bits(rf.epfSharpLut[i]), i < 7 ? ", " : "],\n");
            pw.printf(" \"epf_channel_scale\": [%s, %s, %s], \"epf_pass0_sigma_scale\": %s, \"epf_pass2_sigma_scale\": %s, \"epf_border_sad_mul\": %s,%n",
                bits(rf.epfChannelScale[0]), bits(rf.epfChannelScale[1]), bits(rf.epfChannelScale[2]), bits(rf.epfPass0SigmaScale),
                bits(rf.epfPass2SigmaScale), bits(rf.epfBorderSadMul));
            pw.printf(" \"do_ycbcr\": %d, \"floats_are\": \"IEEE-754 bit patterns (Float.floatToRawIntBits)\"}%n", header.doYCbCr ? 1 : 0);
        } catch (IOException e) {
            throw new IllegalStateException(e);
        }
        flushManifest();
    }

    private static String bits(float f) {
        return Integer.toString(Float.floatToRawIntBits(f));
    }

    /** Frame.decodeFrame, after Gaborish and the edge-preserving filter. */
    public static void afterFilters(Frame frame) {
        if (DIR == null || frame.getFrameHeader().encoding != FrameFlags.VARDCT)
            return;
        planes(frame, "after_filters");
        flushManifest();
    }

    /** JXLCodestreamDecoder.decode, after performColorTransforms: the matrix actually used (primaries folded in) and the planes. */
    public static void afterColor(Frame frame, OpsinInverseMatrix matrix, ImageHeader imageHeader) {
        if (DIR == null || frame.getFrameHeader().encoding != FrameFlags.VARDCT)
            return;
        planes(frame, "after_color");
        if (matrix != null) {
            float[][] m = (float[][])get(matrix, "matrix");
            float[] bias = (float[])get(matrix, "opsinBias");
            try (PrintWriter pw = new PrintWriter(DIR + "/" + String.format("f%d_color.json", frameIndex))) {
                pw.printf("{\"opsin_matrix\": [");
                for (int i = 0; i < 9; i++)
                    pw.printf("%s%s", bits(m[i / 3][i % 3]), i < 8 ? ", " : "], ");
                pw.printf("\"opsin_bias\": [%s, %s, %s], \"intensity_target\": %s}%n", bits(bias[0]), bits(bias[1]), bits(bias[2]),
                    bits(imageHeader.getToneMapping().intensityTarget));
            } catch (IOException e) {
                throw new IllegalStateException(e);
            }
        }
        flushManifest();
    }

    private static void flushManifest() {
        try (PrintWriter pw = new PrintWriter(DIR + "/manifest.json")) {
            String body = manifest.toString().trim();
            if (body.endsWith(","))
                body = body.substring(0, body.length() - 1);
            pw.printf("[%n%s%n]%n", body);
        } catch (IOException e) {
            throw new IllegalStateException(e);
        }
    }
}
