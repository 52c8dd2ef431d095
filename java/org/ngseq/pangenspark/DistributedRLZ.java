/*
 * DistributedRLZ.java -- drop-in main class for `spark-submit --class org.ngseq.pangenspark.DistributedRLZ`
 * (replaces DistributedRLZ.scala of NGSeq/PanGenSpark) that forwards the work to libdrlz.so through the
 * Java 22 foreign function API (java.lang.foreign, "Panama"); no JNI glue, no C headers needed.
 *
 * UNVERIFIED: this image has no JDK, so the file is shipped as source only (see INTEGRATION.md).  It holds no
 * logic beyond argument plumbing and file IO; every computation happens behind include/drlz.h.
 *
 *   java --enable-native-access=ALL-UNNAMED -Ddrlz.lib=/path/to/libdrlz.so \
 *        org.ngseq.pangenspark.DistributedRLZ localradix localref dataGlob fsURI outDir gaplessOut
 *
 * Same six arguments as DistributedRLZ.scala:37-43.  fsURI "" or file:/// writes to the local file system;
 * for hdfs:// wrap the two write calls with org.apache.hadoop.fs.FileSystem exactly as DistributedRLZ.scala:251-257
 * does (left out here so the class compiles without Hadoop on the class path).
 */
package org.ngseq.pangenspark;

import java.io.IOException;
import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.nio.channels.FileChannel;
import java.nio.file.FileSystems;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.PathMatcher;
import java.nio.file.StandardOpenOption;
import java.util.ArrayList;
import java.util.Comparator;
import java.util.List;
import java.util.stream.Stream;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class DistributedRLZ {
    private static final Linker LINKER = Linker.nativeLinker();

    private static MethodHandle fn(SymbolLookup lib, String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(lib.find(name).orElseThrow(), fd);
    }

    public static void main(String[] args) throws Throwable {
        if (args.length < 6) {
            System.err.println("usage: DistributedRLZ localradix localref dataGlob fsURI outDir gaplessOut");
            System.exit(2);
        }
        String localref = args[1], dataGlob = args[2], outDir = args[4], gaplessOut = args[5];
        try (Arena arena = Arena.ofConfined()) {
            SymbolLookup lib = SymbolLookup.libraryLookup(System.getProperty("drlz.lib", "libdrlz.so"), arena);
            MethodHandle create = fn(lib, "drlz_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
            MethodHandle destroy = fn(lib, "drlz_destroy", FunctionDescriptor.ofVoid(ADDRESS));
            MethodHandle lastError = fn(lib, "drlz_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
            MethodHandle buildIndex = fn(lib, "drlz_build_index", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
            MethodHandle parse = fn(lib, "drlz_parse", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT,
                    ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));

            MemorySegment ctxOut = arena.allocate(ADDRESS);
            int device = Integer.parseInt(System.getProperty("drlz.device", "0"));
            if ((int) create.invoke(ctxOut, device) != 0) throw new IllegalStateException("drlz_create failed: no CUDA device (there is no CPU fallback)");
            MemorySegment ctx = ctxOut.get(ADDRESS, 0);

            System.out.println("Create suffix");
            // off-heap mapping: byte[] cannot hold a 3.1 Gbp reference (SURVEY.md 8(b))
            MemorySegment ref;
            long n;
            try (FileChannel ch = FileChannel.open(Path.of(localref), StandardOpenOption.READ)) {
                ref = ch.map(FileChannel.MapMode.READ_ONLY, 0, ch.size(), arena);
                n = ch.size();
                while (n > 0 && (ref.get(java.lang.foreign.ValueLayout.JAVA_BYTE, n - 1) == '\n'
                        || ref.get(java.lang.foreign.ValueLayout.JAVA_BYTE, n - 1) == '\r')) n--;
            }
            check(ctx, lastError, (int) buildIndex.invoke(ctx, ref, n));
            System.out.println("Load suffix\nbroadcasting\nstarted encoding");

            Files.createDirectories(Path.of(outDir));
            Files.createDirectories(Path.of(gaplessOut));
            List<Path> inputs = expand(dataGlob);
            MemorySegment mOut = arena.allocate(JAVA_LONG), nPairs = arena.allocate(JAVA_LONG);
            int index = 0;
            for (Path in : inputs) {
                try (Arena row = Arena.ofConfined(); FileChannel ch = FileChannel.open(in, StandardOpenOption.READ)) {
                    long cols = ch.size();
                    MemorySegment msa = ch.map(FileChannel.MapMode.READ_ONLY, 0, cols, row);
                    MemorySegment gapless = row.allocate(cols + 16);
                    long cap = cols / 32 + 1024;
                    MemorySegment pairs = row.allocate(16 * cap, 16);
                    int rc = (int) parse.invoke(ctx, msa, cols, 1, gapless, mOut, pairs, cap, nPairs);
                    if (rc == -4) {                                   // DRLZ_E_NOSPACE: required count is in *n_pairs
                        cap = nPairs.get(JAVA_LONG, 0) + 16;
                        pairs = row.allocate(16 * cap, 16);
                        rc = (int) parse.invoke(ctx, msa, cols, 1, gapless, mOut, pairs, cap, nPairs);
                    }
                    check(ctx, lastError, rc);
                    String name = in.getFileName().toString();
                    // <outDir>/<basename>.lz: the records are already int64 LE (pos, len) -- DistributedRLZ.scala:263-284
                    try (FileChannel out = FileChannel.open(Path.of(outDir, name + ".lz"), StandardOpenOption.CREATE,
                            StandardOpenOption.WRITE, StandardOpenOption.TRUNCATE_EXISTING)) {
                        out.write(pairs.asSlice(0, 16 * nPairs.get(JAVA_LONG, 0)).asByteBuffer());
                    }
                    // gapless FASTA record -- DistributedRLZ.scala:75-80
                    try (FileChannel out = FileChannel.open(Path.of(gaplessOut, String.format("part-%05d", index)),
                            StandardOpenOption.CREATE, StandardOpenOption.WRITE, StandardOpenOption.TRUNCATE_EXISTING)) {
                        out.write(java.nio.ByteBuffer.wrap((">" + name + "\n").getBytes()));
                        out.write(gapless.asSlice(0, mOut.get(JAVA_LONG, 0)).asByteBuffer());
                        out.write(java.nio.ByteBuffer.wrap("\n".getBytes()));
                    }
                }
                index++;
            }
            Files.write(Path.of(gaplessOut, "_SUCCESS"), new byte[0]);
            destroy.invoke(ctx);
        }
    }

    private static void check(MemorySegment ctx, MethodHandle lastError, int rc) throws Throwable {
        if (rc == 0) return;
        MemorySegment msg = ((MemorySegment) lastError.invoke(ctx)).reinterpret(4096);
        throw new IOException("libdrlz error " + rc + ": " + msg.getString(0));
    }

    /** Sorted by file name: the order `hdfs dfs -getmerge outDir/*.NN.lz` concatenates (index_chr.sh:16). */
    private static List<Path> expand(String glob) throws IOException {
        String g = glob.startsWith("file://") ? glob.substring(7) : glob;
        Path dir = Path.of(g).getParent() == null ? Path.of(".") : Path.of(g).getParent();
        PathMatcher m = FileSystems.getDefault().getPathMatcher("glob:" + Path.of(g).getFileName());
        List<Path> out = new ArrayList<>();
        try (Stream<Path> s = Files.list(dir)) {
            s.filter(p -> m.matches(p.getFileName())).forEach(out::add);
        }
        out.sort(Comparator.comparing(p -> p.getFileName().toString()));
        return out;
    }
}
