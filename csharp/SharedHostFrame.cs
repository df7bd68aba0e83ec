// SharedHostFrame.cs -- the reassembled lit buffer of one node in HOST memory, written by every rank directly.
//
// SOURCE ONLY (no .NET toolchain in this image); the C# twin of illuminant_b200/sharding.py::SharedHostFrame, which is the
// implementation bench.py and tests/test_multirank_gloo.py run.  One process per GPU: rank 0 creates a memory-mapped file
// (on Linux under /dev/shm, i.e. POSIX shared memory), every rank opens it, page-locks its view with ilb_host_register and
// passes Rows(r0) as `lightmapOut` of its own ilb_render_lighting_frame call with IlbLightingFrame.RowBegin / RowEnd = its
// band.  Each rank's rows then travel from its GPU into this frame over the rank's own PCIe link: no collective, no GPU
// barrier.  Layout: a 4096-byte header (int64 slot 8 * k = sequence number of rank k, slot 8 * world = the consumer's),
// then `depth` frames of height * width texels of HalfVector4, each starting on a page (frame s lives in slot s % depth: with
// two slots the ranks work on frame s + 1 while the consumer still holds frame s).
using System;
using System.IO;
using System.IO.MemoryMappedFiles;
using System.Threading;

namespace Squared.Illuminant.Native {
    public sealed unsafe class SharedHostFrame : IDisposable {
        public const int HeaderBytes = 4096;
        public readonly int Rank, World, Width, Height, TexelBytes, Depth;
        readonly long FrameBytes;
        readonly IntPtr Context;
        readonly MemoryMappedFile File;
        readonly MemoryMappedViewAccessor View;
        readonly byte* Base;
        readonly string Path;

        public SharedHostFrame (IntPtr ctx, string name, int width, int height, int rank, int world, int texelBytes = 8, int depth = 1) {
            if (8 * (world + 1) * 8 > HeaderBytes)
                throw new ArgumentOutOfRangeException(nameof(world));
            Context = ctx; Rank = rank; World = world; Width = width; Height = height; TexelBytes = texelBytes; Depth = Math.Max(depth, 1);
            FrameBytes = ((long)width * height * texelBytes + 4095) / 4096 * 4096;
            long bytes = HeaderBytes + Depth * FrameBytes;
            Path = System.IO.Path.Combine("/dev/shm", name);
            if (rank == 0) {
                // created under a temporary name and renamed, so that it appears at full size, zero-filled
                using (var fs = new FileStream(Path + ".tmp", FileMode.Create, FileAccess.ReadWrite, FileShare.ReadWrite))
                    fs.SetLength(bytes);
                System.IO.File.Delete(Path);
                System.IO.File.Move(Path + ".tmp", Path);
            } else {
                for (int tries = 0; !(System.IO.File.Exists(Path) && new FileInfo(Path).Length == bytes); tries++) {
                    if (tries > 2000)
                        throw new TimeoutException("shared host frame " + Path + " did not appear");
                    Thread.Sleep(5);
                }
            }
            File = MemoryMappedFile.CreateFromFile(Path, FileMode.Open, null, bytes, MemoryMappedFileAccess.ReadWrite);
            View = File.CreateViewAccessor(0, bytes, MemoryMappedFileAccess.ReadWrite);
            byte* p = null;
            View.SafeMemoryMappedViewHandle.AcquirePointer(ref p);
            Base = p;
            B200.Check(ctx, B200.ilb_host_register(ctx, Base, (UIntPtr)(ulong)bytes));
        }

        /// <summary>What a rank passes as lightmapOut for the band of frame s that starts at row rowBegin.</summary>
        public void* Rows (int rowBegin, long s = 0) => Base + HeaderBytes + (s % Depth) * FrameBytes + (long)rowBegin * Width * TexelBytes;

        long* Slot (int index) => (long*)(Base + 64 * index);

        static void Spin (Func<bool> done, string what, int timeoutMs = 60000) {
            var sw = new SpinWait();
            long deadline = Environment.TickCount64 + timeoutMs;
            while (!done()) {
                sw.SpinOnce();
                if (Environment.TickCount64 > deadline)
                    throw new TimeoutException(what);
            }
        }

        /// <summary>Blocks until frame s - Depth has been taken by the consumer: its slot is about to be overwritten.</summary>
        public void Begin (long s) => Spin(() => Volatile.Read(ref *Slot(World)) >= s - Depth, "frame " + (s - Depth) + " was never released");
        /// <summary>This rank's band of frame s is in place (call after ilb_render_lighting_frame has returned).</summary>
        public void Publish (long s) => Volatile.Write(ref *Slot(Rank), s);
        /// <summary>Consumer: blocks until every rank has published frame s.</summary>
        public void WaitComplete (long s) => Spin(() => {
            for (int k = 0; k < World; k++)
                if (Volatile.Read(ref *Slot(k)) < s)
                    return false;
            return true;
        }, "frame " + s + " incomplete");
        /// <summary>Consumer: frame s has been taken.</summary>
        public void Release (long s) => Volatile.Write(ref *Slot(World), s);

        public void Dispose () {
            B200.ilb_host_unregister(Context, Base);
            View.SafeMemoryMappedViewHandle.ReleasePointer();
            View.Dispose();
            File.Dispose();
            if (Rank == 0)
                System.IO.File.Delete(Path);
        }
    }
}
