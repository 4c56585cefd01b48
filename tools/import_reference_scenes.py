"""Import the reference's committed *data fixtures* (serde scene dumps and texture images) into scenes/.

Run once in the build container (where /root/reference is mounted); outputs are committed because
/root/reference does not exist on the GPU box.  Scene dumps are stored gzip-compressed, byte-identical
after decompression.  No reference source code is copied.
"""
import gzip, os, shutil, sys
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")
os.makedirs(os.path.join(OUT, "assets"), exist_ok=True)
for name in ("suzanne.yml", "teapot.yml", "conics.yml"):
    with open(os.path.join(REF, "scenes", name), "rb") as f, gzip.GzipFile(os.path.join(OUT, name + ".gz"), "wb", mtime=0) as g:
        g.write(f.read())
for name in ("earthmap.jpg", "uvmap.png"):
    shutil.copyfile(os.path.join(REF, name), os.path.join(OUT, "assets", name))
print("imported")
