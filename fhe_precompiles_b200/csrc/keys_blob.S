/* Network key pair linked into the library, the equivalent of the reference's
 * include_bytes!("data/network.pub") / ("data/network.pri") at /root/reference/src/fhe.rs:118-119. */
    .section .rodata
    .balign 16
    .global fhe_b200_network_pub
    .global fhe_b200_network_pub_end
fhe_b200_network_pub:
    .incbin NETWORK_PUB_PATH
fhe_b200_network_pub_end:
    .balign 16
    .global fhe_b200_network_pri
    .global fhe_b200_network_pri_end
fhe_b200_network_pri:
    .incbin NETWORK_PRI_PATH
fhe_b200_network_pri_end:
    .section .note.GNU-stack,"",@progbits
