"""Import-path shim: the reference does `import pointnet2._ext as _ext`
(detection/Votenet/pointnet2/pointnet2_utils.py:25-33).  With the repo root on sys.path this
package satisfies that import with the B200-native ops, so the reference's pointnet2_utils.py /
pointnet2_modules.py run unchanged on top of libb2r.so.  See INTEGRATION.md."""
