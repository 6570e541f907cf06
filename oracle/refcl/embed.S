/* Embeds the reference's OpenCL kernel sources, from where they lie under /root/reference/kernels
 * (-I on the assembler command line, see ../Makefile), into oracle/_ref/librefcl.so.  The
 * reference's own build does the same with utils/clc2cpp.py.  Nothing is copied into the repo. */
    .section .rodata
    .macro embed sym, file
    .global \sym
    .hidden \sym
\sym:
    .incbin "\file"
    .byte 0
    .endm
    embed refcl_src_octree, octree.cl
    embed refcl_src_mls, mls.cl
    embed refcl_src_marching, marching.cl
    embed refcl_src_scale_bias, scale_bias.cl
    .section .note.GNU-stack,"",@progbits
