/*
 * HdfsSink.java -- the HDFS side of the drop-in main class (DistributedRLZ.java): outputs written through Hadoop's
 * FileSystem API exactly as the reference does at DistributedRLZ.scala:251-257
 *     val fs = FileSystem.get(new URI(hdfsurl), new Configuration()); val fos = fs.create(new Path(out + "/" + name + ".lz"))
 * The only class that needs Hadoop on the class path (spark-submit provides it; standalone: `hadoop classpath`); loaded
 * by name, and only when fsURI starts with hdfs://.
 *
 * UNVERIFIED: this image has no JDK and no Hadoop; shipped as source only (see INTEGRATION.md).
 */
package org.ngseq.pangenspark;

import java.io.IOException;
import java.io.OutputStream;
import java.net.URI;

import java.nio.channels.Channels;
import java.nio.channels.ReadableByteChannel;
import java.util.ArrayList;
import java.util.List;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileStatus;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;

public final class HdfsSink implements DistributedRLZ.Sink {
    private final FileSystem fs;

    public HdfsSink(String fsUri) throws IOException {
        fs = FileSystem.get(URI.create(fsUri), new Configuration());
    }

    @Override
    public void mkdirs(String dir) throws IOException {
        fs.mkdirs(new Path(dir));
    }

    @Override
    public OutputStream create(String path) throws IOException {
        return fs.create(new Path(path), true);          // overwrite, as fs.create(path) does
    }

    @Override
    public List<String[]> glob(String pattern) throws IOException {
        List<String[]> out = new ArrayList<>();
        FileStatus[] st = fs.globStatus(new Path(pattern));
        if (st != null) for (FileStatus f : st) if (f.isFile()) out.add(new String[] {f.getPath().toString(), Long.toString(f.getLen())});
        return out;
    }

    @Override
    public ReadableByteChannel open(String path) throws IOException {
        return Channels.newChannel(fs.open(new Path(path)));
    }
}
