/* Embeds data/lab_lut_s16.bin (OpenCV float RGB->Lab lattice, tools/gen_lab_lut.py) into libf3ps.so. */
    .section .rodata
    .balign 64
    .global f3ps_lab_lut_begin
    .global f3ps_lab_lut_end
f3ps_lab_lut_begin:
    .incbin "lab_lut_s16.bin"
f3ps_lab_lut_end:
    .byte 0
    .section .note.GNU-stack,"",@progbits
