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
 * Same six arguments as DistributedRLZ.scala:37-43.  fsURI "" or file:/// writes to the local file system; with
 * hdfs://... the outputs go through org.apache.hadoop.fs.FileSystem exactly as DistributedRLZ.scala:251-257 does
 * (HdfsSink.java next to this file; it is the only class that needs Hadoop on the class path, and it is only
 * loaded for an hdfs:// URI).
 *
 * Same streaming shape as the native CLI (pangenspark_b200/cli/drlz_main.cpp): the rows of a batch are read into one
 * pinned slot (drlz_host_alloc) by a pool of reader threads, the batch is handed to the library with drlz_submit (the
 * library copies and parses batch k + 1 while batch k is being written), drlz_wait collects the batches in order and the
 * same pool writes <basename>.lz and the gapless FASTA records from the pinned result buffers.
 * -Ddrlz.slots (3), -Ddrlz.batchBytes (128 MiB), -Ddrlz.ioThreads (8), -Ddrlz.device (0).
 */
package org.ngseq.pangenspark;

import java.io.IOException;
import java.io.OutputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.nio.ByteBuffer;
import java.nio.channels.Channels;
import java.nio.channels.FileChannel;
import java.nio.channels.WritableByteChannel;
import java.nio.file.FileSystems;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.PathMatcher;
import java.nio.file.StandardOpenOption;
import java.util.ArrayDeque;
import java.util.ArrayList;
import java.util.Comparator;
import java.util.List;
import java.util.concurrent.ExecutorService;
import java.util.concurrent.Executors;
import java.util.concurrent.Future;
import java.util.stream.Stream;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class DistributedRLZ {
    private static final Linker LINKER = Linker.nativeLinker();

    private static MethodHandle fn(SymbolLookup lib, String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(lib.find(name).orElseThrow(), fd);
    }

    /** Where the outputs go: the local file system, or HDFS through Hadoop's FileSystem (HdfsSink). */
    public interface Sink {
        void mkdirs(String dir) throws IOException;
        OutputStream create(String path) throws IOException;        // overwrites, like fs.create(path) at DistributedRLZ.scala:257
        /** {path, size} of every file the glob matches (Spark reads dataGlob from the same file system, drlz.sh:15). */
        List<String[]> glob(String pattern) throws IOException;
        java.nio.channels.ReadableByteChannel open(String path) throws IOException;
    }

    static final class LocalSink implements Sink {
        public void mkdirs(String dir) throws IOException { Files.createDirectories(Path.of(strip(dir))); }
        public OutputStream create(String path) throws IOException {
            return Files.newOutputStream(Path.of(strip(path)), StandardOpenOption.CREATE, StandardOpenOption.WRITE, StandardOpenOption.TRUNCATE_EXISTING);
        }
        public List<String[]> glob(String pattern) throws IOException {
            List<String[]> out = new ArrayList<>();
            for (Path p : expand(strip(pattern))) out.add(new String[] {p.toString(), Long.toString(Files.size(p))});
            return out;
        }
        public java.nio.channels.ReadableByteChannel open(String path) throws IOException { return FileChannel.open(Path.of(strip(path)), StandardOpenOption.READ); }
        private static String strip(String p) { return p.startsWith("file://") ? p.substring(7) : p; }
    }

    /** One batch in flight: a pinned input buffer, pinned result buffers, the arrays the C ABI reads and fills. */
    private static final class Slot {
        MemorySegment in, gapless, pairs;          // pinned (drlz_host_alloc), sized for one batch
        MemorySegment rowPtrs, glPtrs, cols, m, np, ticket;
        long pairsCap;
        List<String[]> files = new ArrayList<>();          // {path, size}
        long[] off;
    }

    public static void main(String[] args) throws Throwable {
        if (args.length < 6) {
            System.err.println("usage: DistributedRLZ localradix localref dataGlob fsURI outDir gaplessOut");
            System.exit(2);
        }
        final String localref = args[1], dataGlob = args[2], fsUri = args[3], outDir = args[4], gaplessOut = args[5];
        final int nslots = Integer.getInteger("drlz.slots", 3), ioThreads = Integer.getInteger("drlz.ioThreads", 8);
        final long batchBytes = Long.getLong("drlz.batchBytes", 128L << 20);
        final Sink sink = fsUri.startsWith("hdfs://")
                ? (Sink) Class.forName("org.ngseq.pangenspark.HdfsSink").getConstructor(String.class).newInstance(fsUri)
                : new LocalSink();
        try (Arena arena = Arena.ofShared()) {
            SymbolLookup lib = SymbolLookup.libraryLookup(System.getProperty("drlz.lib", "libdrlz.so"), arena);
            MethodHandle create = fn(lib, "drlz_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
            MethodHandle destroy = fn(lib, "drlz_destroy", FunctionDescriptor.ofVoid(ADDRESS));
            MethodHandle lastError = fn(lib, "drlz_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
            MethodHandle buildIndex = fn(lib, "drlz_build_index", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
            MethodHandle hostAlloc = fn(lib, "drlz_host_alloc", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
            MethodHandle hostFree = fn(lib, "drlz_host_free", FunctionDescriptor.ofVoid(ADDRESS));
            // int drlz_submit(ctx, rows**, cols*, H, strip, gapless**, m*, pairs*, cap, n_pairs*, ticket*)
            MethodHandle submit = fn(lib, "drlz_submit", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT,
                    ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
            MethodHandle waitFor = fn(lib, "drlz_wait", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
            MethodHandle parseBatch = fn(lib, "drlz_parse_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT,
                    ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));

            MemorySegment ctxOut = arena.allocate(ADDRESS);
            int device = Integer.getInteger("drlz.device", 0);
            if ((int) create.invoke(ctxOut, device) != 0) throw new IllegalStateException("drlz_create failed: no CUDA device (there is no CPU fallback)");
            final MemorySegment ctx = ctxOut.get(ADDRESS, 0);

            System.out.println("Create suffix");
            // off-heap mapping: byte[] cannot hold a 3.1 Gbp reference (SURVEY.md 8(b))
            try (FileChannel ch = FileChannel.open(Path.of(localref), StandardOpenOption.READ)) {
                MemorySegment ref = ch.map(FileChannel.MapMode.READ_ONLY, 0, ch.size(), arena);
                long n = ch.size();
                while (n > 0 && (ref.get(JAVA_BYTE, n - 1) == '\n' || ref.get(JAVA_BYTE, n - 1) == '\r')) n--;
                check(ctx, lastError, (int) buildIndex.invoke(ctx, ref, n));
            }
            System.out.println("Load suffix\nbroadcasting\nstarted encoding");

            List<String[]> inputs = sink.glob(dataGlob);          // sorted by file name
            inputs.sort(Comparator.comparing(f -> baseName(f[0])));
            if (inputs.isEmpty()) throw new IOException("Path does not exist: " + dataGlob);      // what Spark says
            sink.mkdirs(outDir);
            sink.mkdirs(gaplessOut);
            // batches of consecutive files, at most batchBytes each
            List<List<String[]>> batches = new ArrayList<>();
            long slotBytes = 0; int slotRows = 1;
            {
                List<String[]> cur = new ArrayList<>(); long acc = 0, padded = 0;
                for (String[] p : inputs) {
                    long sz = Long.parseLong(p[1]);
                    if (!cur.isEmpty() && acc + sz > batchBytes) { batches.add(cur); cur = new ArrayList<>(); acc = 0; padded = 0; }
                    cur.add(p); acc += sz; padded += ((sz + 15) & ~15L) + 16;
                    slotBytes = Math.max(slotBytes, padded + 16); slotRows = Math.max(slotRows, cur.size());
                }
                if (!cur.isEmpty()) batches.add(cur);
            }
            ExecutorService pool = Executors.newFixedThreadPool(ioThreads);
            ArrayDeque<Slot> free = new ArrayDeque<>(), inGpu = new ArrayDeque<>();
            List<Slot> all = new ArrayList<>();
            for (int i = 0; i < Math.min(nslots, batches.size() + 1); i++) {
                Slot s = new Slot();
                s.in = ((MemorySegment) hostAlloc.invoke(slotBytes)).reinterpret(slotBytes);
                s.gapless = ((MemorySegment) hostAlloc.invoke(slotBytes)).reinterpret(slotBytes);
                s.pairsCap = slotBytes / 64 + 1024L * slotRows;
                s.pairs = ((MemorySegment) hostAlloc.invoke(16 * s.pairsCap)).reinterpret(16 * s.pairsCap);
                s.rowPtrs = arena.allocate(ADDRESS, slotRows); s.glPtrs = arena.allocate(ADDRESS, slotRows);
                s.cols = arena.allocate(JAVA_LONG, slotRows); s.m = arena.allocate(JAVA_LONG, slotRows);
                s.np = arena.allocate(JAVA_LONG, slotRows); s.ticket = arena.allocate(JAVA_LONG);
                free.add(s); all.add(s);
            }
            int firstIndex = 0, nextBatch = 0;
            int[] batchFirst = new int[batches.size()];
            for (int b = 0; b < batches.size(); b++) { batchFirst[b] = firstIndex; firstIndex += batches.get(b).size(); }
            List<Future<?>> writes = new ArrayList<>();
            int collected = 0;
            while (collected < batches.size()) {
                // fill and submit while a slot is free
                while (nextBatch < batches.size() && !free.isEmpty()) {
                    Slot s = free.poll();
                    s.files = batches.get(nextBatch);
                    int H = s.files.size();
                    s.off = new long[H];
                    long at = 0;
                    List<Future<?>> reads = new ArrayList<>();
                    for (int h = 0; h < H; h++) {
                        final long sz = Long.parseLong(s.files.get(h)[1]), o = at;
                        final String p = s.files.get(h)[0];
                        s.off[h] = o; at += ((sz + 15) & ~15L) + 16;
                        reads.add(pool.submit(() -> {          // straight into the pinned slot
                            try (java.nio.channels.ReadableByteChannel ch = sink.open(p)) {
                                long got = 0;
                                while (got < sz) {             // a ByteBuffer view holds at most 2 GiB
                                    ByteBuffer dst = s.in.asSlice(o + got, Math.min(sz - got, 1L << 30)).asByteBuffer();
                                    while (dst.hasRemaining()) { int r = ch.read(dst); if (r < 0) throw new IOException("short read: " + p); got += r; }
                                }
                            }
                            return null;
                        }));
                        s.rowPtrs.setAtIndex(ADDRESS, h, s.in.asSlice(o));
                        s.glPtrs.setAtIndex(ADDRESS, h, s.gapless.asSlice(o));
                        s.cols.setAtIndex(JAVA_LONG, h, sz);           // the library drops the line terminator itself
                    }
                    for (Future<?> f : reads) f.get();
                    check(ctx, lastError, (int) submit.invoke(ctx, s.rowPtrs, s.cols, H, 1, s.glPtrs, s.m, s.pairs, s.pairsCap, s.np, s.ticket));
                    s.ticket.set(JAVA_LONG, 0, s.ticket.get(JAVA_LONG, 0));
                    inGpu.add(s);
                    nextBatch++;
                }
                // collect the oldest batch, hand its results to the writers
                final Slot s = inGpu.poll();
                int rc = (int) waitFor.invoke(ctx, s.ticket.get(JAVA_LONG, 0));
                final int H = s.files.size();
                if (rc == -4) {                                       // DRLZ_E_NOSPACE: the required counts are in n_pairs
                    long need = 16;
                    for (int h = 0; h < H; h++) need += s.np.getAtIndex(JAVA_LONG, h);
                    hostFree.invoke(s.pairs);
                    s.pairsCap = need + need / 8;
                    s.pairs = ((MemorySegment) hostAlloc.invoke(16 * s.pairsCap)).reinterpret(16 * s.pairsCap);
                    rc = (int) parseBatch.invoke(ctx, s.rowPtrs, s.cols, H, 1, s.glPtrs, s.m, s.pairs, s.pairsCap, s.np);
                }
                check(ctx, lastError, rc);
                final int base = batchFirst[collected];
                List<Future<?>> mine = new ArrayList<>();
                long poff = 0;
                for (int h = 0; h < H; h++) {
                    final int hh = h; final long po = poff, cnt = s.np.getAtIndex(JAVA_LONG, h), mh = s.m.getAtIndex(JAVA_LONG, h);
                    poff += cnt;
                    mine.add(pool.submit(() -> {
                        String name = baseName(s.files.get(hh)[0]);
                        // <outDir>/<basename>.lz: the records are already int64 LE (pos, len) -- DistributedRLZ.scala:263-284
                        try (WritableByteChannel out = Channels.newChannel(sink.create(outDir + "/" + name + ".lz"))) {
                            writeAll(out, s.pairs.asSlice(16 * po, 16 * cnt));
                        }
                        // gapless FASTA record, one part file per record in list order -- DistributedRLZ.scala:75-80
                        try (WritableByteChannel out = Channels.newChannel(sink.create(gaplessOut + "/" + String.format("part-%05d", base + hh)))) {
                            out.write(ByteBuffer.wrap((">" + name + "\n").getBytes()));
                            writeAll(out, s.gapless.asSlice(s.off[hh], mh));
                            out.write(ByteBuffer.wrap("\n".getBytes()));
                        }
                        return null;
                    }));
                }
                // the slot is free again once its files are written (the reads of the next batch overlap with them)
                writes.add(pool.submit(() -> { for (Future<?> f : mine) f.get(); synchronized (free) { free.add(s); free.notifyAll(); } return null; }));
                collected++;
                if (nextBatch < batches.size()) synchronized (free) { while (free.isEmpty()) free.wait(); }
            }
            for (Future<?> f : writes) f.get();
            pool.shutdown();
            try (OutputStream ok = sink.create(gaplessOut + "/_SUCCESS")) { ok.flush(); }
            for (Slot s : all) { hostFree.invoke(s.in); hostFree.invoke(s.gapless); hostFree.invoke(s.pairs); }
            destroy.invoke(ctx);
        }
    }

    private static String baseName(String path) { return path.substring(path.lastIndexOf('/') + 1); }

    private static void writeAll(WritableByteChannel out, MemorySegment seg) throws IOException {
        long at = 0, n = seg.byteSize();
        while (at < n) {                                             // a ByteBuffer view holds at most 2 GiB
            ByteBuffer b = seg.asSlice(at, Math.min(n - at, 1L << 30)).asByteBuffer();
            at += b.remaining();
            while (b.hasRemaining()) out.write(b);
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
        if (!Files.isDirectory(dir)) return out;
        try (Stream<Path> s = Files.list(dir)) {
            s.filter(p -> m.matches(p.getFileName())).forEach(out::add);
        }
        out.sort(Comparator.comparing(p -> p.getFileName().toString()));
        return out;
    }
}
